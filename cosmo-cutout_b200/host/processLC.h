/*
 * host/processLC.h -- the two cutout entry points of the reference (src/processLC.h:41-48), same names, argument
 * order and meaning, re-implemented over the C ABI of include/cosmo_cutout_b200.h.
 *
 * Differences a caller can observe are listed in INTEGRATION.md; the on-disk results (one raw little-endian file
 * per column under <out><prefix>Cutout<step>/, properties.csv) are byte-identical to a single-rank run of the
 * reference.  `myrank`/`numranks` keep their position: this implementation is one host process driving
 * `numranks` GPUs (rank r of the reference's output layout = GPU r's contiguous index shard), so it must be
 * called with myrank == 0.
 */
#ifndef CC_HOST_PROCESSLC_H
#define CC_HOST_PROCESSLC_H

#include <string>
#include <vector>

#define PI 3.14159265 /* reference processLC.h:35 */
#define ARCSEC 3600.0 /* reference processLC.h:36 */

/* use case 1: constant angular bounds {lo, hi} in arcsec (reference processLC.cpp:15-429) */
void processLC(std::string dir_name, std::string out_dir, std::vector<std::string> step_strings,
               std::vector<float> theta_bounds, std::vector<float> phi_bounds, int myrank, int numranks,
               bool verbose, bool timeit, bool overwrite, bool positionOnly);

/* use case 2: halo-centred square cutouts of `boxLength` arcmin (reference processLC.cpp:440-1757) */
void processLC(std::string dir_name, std::vector<std::string> out_dirs, std::vector<std::string> step_strings,
               std::vector<float> halo_pos, std::vector<float> halo_props, float boxLength, int myrank,
               int numranks, bool verbose, bool timeit, bool overwrite, bool positionOnly,
               bool forceWriteProps, bool propsOnly);

namespace cchost {
/* Extension (no counterpart in the reference): keep every lightcone step a processLC call reads resident in HBM, so
 * that later calls in the same process -- another boxLength, another window -- neither read nor upload it again.
 * Off by default (one call cuts each step once and then streams only x,y,z); lc_cutout turns it on when several
 * --boxLength values are given; CC_RESIDENT_STEPS=1 in the environment does the same. */
void resident_steps_enable(bool on);
/* Extension: the cutouts of SEVERAL processLC calls that differ only in boxLength (and output directories) as one
 * call: entry b of out_dirs_of_box lists one directory per halo for boxLengths[b].  Every (halo, box length) pair is an
 * independent cutout -- files identical to those of the separate calls -- but each lightcone step is read, uploaded and
 * scanned ONCE for all of them (the device pass takes per-halo bounds; the reference's one-boxLength-per-call signature,
 * processLC.h:46, is why BASELINE config 5 otherwise needs a pass per size). */
void processLC_boxes(const std::string& dir_name, const std::vector<std::vector<std::string>>& out_dirs_of_box,
                     const std::vector<std::string>& step_strings, const std::vector<float>& halo_pos,
                     const std::vector<float>& halo_props, const std::vector<float>& boxLengths, int myrank, int numranks,
                     bool verbose, bool timeit, bool overwrite, bool positionOnly, bool forceWriteProps, bool propsOnly);
/* destroy the GPU contexts (call before main returns; never from a static destructor) */
void shutdown_gpus();
}  // namespace cchost

#endif
