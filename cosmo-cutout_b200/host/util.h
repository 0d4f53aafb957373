/*
 * host/util.h -- host utilities of the lc_cutout driver: lightcone directory discovery, step selection, output
 * directory preparation, halo list parsing and the step<->redshift maps.  Behaviour-compatible re-write of the
 * reference's src/util.h:96-153 / src/util.cpp:116-552 (the linear algebra of util.cpp:574-902 lives in
 * csrc/host_setup.cpp behind cc_halo_setup; the per-particle parts are device code).
 */
#ifndef CC_HOST_UTIL_H
#define CC_HOST_UTIL_H

#include <stdint.h>

#include <string>
#include <vector>

namespace cchost {

/* fatal error: message to stdout, exit non-zero (the reference calls MPI_Abort(MPI_COMM_WORLD, 0)) */
[[noreturn]] void die(const std::string& msg);

/* util.cpp:116-125 */
bool does_file_exist(const std::string& filename);
/* util.cpp:131-217; massDef "sod": 10 columns per row, "fof": 7.  Unlike the reference,
 * "fof" rows yield exactly 3 properties. */
void readHaloFile(const std::string& haloFileName, std::vector<float>& haloPos, std::vector<std::string>& haloTags,
                  std::vector<float>& haloProps, const std::string& massDef);
/* util.cpp:223-252: names under dir containing "lc" or "STEP" */
int getLCSubdirs(const std::string& dir, std::vector<std::string>& subdirs);
/* util.cpp:258-311: the single un-hashed header file in a step directory */
int getLCFile(const std::string& dir, std::string& file);
/* util.cpp:317-373: steps present with minStep <= step <= maxStep, ascending */
int getLCSteps(int maxStep, int minStep, const std::string& dir, std::vector<std::string>& step_strings);
/* util.cpp:379-479, without its double closedir and examining every entry (Q13).
 * returns 0 ok / 1 non-.bin files present with overwrite (fatal) / 2 non-empty without overwrite (skip halo) */
int prepStepSubdir(const std::string& subdir_name, bool overwrite, bool print, bool verbose);
/* prefix of the step sub-directories = characters before the first digit (processLC.cpp:41-47) */
std::string subdirPrefixOf(const std::vector<std::string>& subdirs);

/* util.cpp:492-552 (defaults util.h:149-153) */
float aToZ(float a);
int zToStep(float z, int totSteps = 499, float maxZ = 200.0f);
float stepToZ(float step, int totSteps = 499, float maxZ = 200.0f);

double wall_seconds();

}  // namespace cchost
#endif
