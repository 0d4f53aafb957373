/*
 * host/main.cpp -- the lc_cutout command line of the reference (src/main.cpp:36-403), same positional arguments
 * and flags, over the B200 cutout path.
 *
 *   lc_cutout <input lightcone dir> <output dir> <min redshift | max step> <max redshift | min step>
 *             { --theta <center> <half> --phi <center> <half>          (use case 1, degrees)
 *             | --halo <x> <y> <z> --boxLength <arcmin>                 (use case 2, one halo)
 *             | --haloFile <file> --boxLength <arcmin> [--massDef sod|fof] }  (use case 2, many halos)
 *             [--verbose] [--timeit] [--overwrite] [--posOnly] [--forceWriteProps] [--propsOnly]
 *
 * Extension: --boxLength takes one value like the reference, or SEVERAL (--boxLength 5 10 20 30).  With several, the
 * results are what separate invocations with each value would write -- the reference's processLC has one boxLength
 * per call (processLC.h:46) -- under <output dir>/boxLength_<value>/, but every lightcone step is read, uploaded and
 * scanned ONCE for all sizes (cchost::processLC_boxes: the device pass takes per-halo bounds).  CC_BOX_PASSES=separate
 * runs one processLC call per size instead, with the steps of the first call kept resident in HBM for the others.
 *
 * Where the reference is started under mpirun with one rank per core, this is ONE process that drives every
 * visible GPU (or CC_NGPUS of them); "ranks" in its log lines are GPUs.
 */
#include <dirent.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <iostream>
#include <string>
#include <vector>

#include "cosmo_cutout_b200.h"
#include "processLC.h"
#include "util.h"

using std::cout;
using std::endl;
using std::string;
using std::vector;
using namespace cchost;

static bool has_flag(const vector<string>& args, const char* shortf, const char* longf) {
    return std::find(args.begin(), args.end(), shortf) != args.end() ||
           std::find(args.begin(), args.end(), longf) != args.end();
}

int main(int argc, char** argv) {
    if (argc < 5) {
        cout << "usage: lc_cutout <input lightcone dir> <output dir> <min redshift> <max redshift> "
                "{--theta c d --phi c d | --halo x y z --boxLength b | --haloFile f --boxLength b} [options]" << endl;
        return EXIT_FAILURE;
    }
    int numranks = cc_device_count();
    if (const char* e = getenv("CC_NGPUS")) numranks = std::max(1, std::min(numranks, atoi(e)));
    const int myrank = 0;
    cout << "\n\n---------- Starting on " << numranks << " GPU ranks ----------" << endl;
    const char cart[3] = {'x', 'y', 'z'};

    string input_lc_dir(argv[1]), out_dir(argv[2]);
    if (input_lc_dir.empty() || out_dir.empty()) die("\nempty directory argument");
    if (input_lc_dir[input_lc_dir.size() - 1] != '/') input_lc_dir += "/"; /* main.cpp:149-159 */
    if (out_dir[out_dir.size() - 1] != '/') out_dir += "/";
    struct stat info;
    if (stat(out_dir.c_str(), &info) != 0) die("\nCannot access specified output directory " + out_dir);
    cout << "Using lightcone at " << input_lc_dir << endl;
    cout << "Writing at " << out_dir << endl;

    /* both bounds above 20 => snapshot numbers, otherwise redshifts (main.cpp:175-193) */
    const float minBound = atof(argv[3]), maxBound = atof(argv[4]);
    int minStep, maxStep;
    float minZ, maxZ;
    if (minBound > 20 && maxBound > 20) {
        maxStep = int(minBound);
        minStep = int(maxBound);
        maxZ = stepToZ(minStep);
        minZ = stepToZ(maxStep);
    } else {
        minZ = minBound;
        maxZ = maxBound;
        maxStep = zToStep(minZ);
        minStep = zToStep(maxZ);
    }
    vector<string> step_strings;
    getLCSteps(maxStep, minStep, input_lc_dir, step_strings);
    cout << "MAX STEP: " << maxStep << endl;
    cout << "MIN STEP: " << minStep << endl;
    cout << "steps to include from z= " << minZ << " to z=" << maxZ << ": " << endl;
    for (const string& s : step_strings) cout << s << " ";
    cout << endl;

    vector<float> theta_cut(2), phi_cut(2), haloPos, haloProps;
    vector<string> haloTags;
    float boxLength = 0;
    vector<float> boxLengths;
    vector<string> boxLabels;
    bool verbose = false, timeit = false, overwrite = false, positionOnly = false, forceWriteProps = false, propsOnly = false;
    string massDef = "sod";

    const vector<string> args(argv + 1, argv + argc);
    const bool customThetaBounds = has_flag(args, "-t", "--theta");
    const bool customPhiBounds = has_flag(args, "-p", "--phi");
    const bool customHalo = has_flag(args, "-h", "--halo");
    const bool customHaloFile = has_flag(args, "-f", "--haloFile");
    const bool customBox = has_flag(args, "-b", "--boxLength");
    const bool customMassDef = has_flag(args, "-m", "--massDef");
    /* main.cpp:238-258: the two use cases must not be mixed */
    if ((customHalo || customHaloFile) ^ customBox) die("\n-h (or -f) and -b options must accompany eachother");
    if (customMassDef && !customHaloFile) die("\n-m does nothing if not used along with -f");
    if (customThetaBounds ^ customPhiBounds) die("\n-t and -p options must accompany eachother");
    if ((customHalo && customThetaBounds) || (customHaloFile && customThetaBounds))
        die("\n-t and -p options must not be used in the case that -h (or -f) and -b arguments are passed");
    if (!customThetaBounds && !customPhiBounds && !customHalo && !customHaloFile && !customBox)
        die("\nValid options are -h, -f, -b, -t, -p, -v, -m, and --timeit. See github readme for help");

    auto need = [&](int i, int k) {
        if (i + k >= argc) die(string("\noption ") + argv[i] + " needs " + std::to_string(k) + " value(s)");
    };
    for (int i = 5; i < argc; ++i) { /* main.cpp:262-317 */
        const char* a = argv[i];
        if (!strcmp(a, "-t") || !strcmp(a, "--theta")) {
            need(i, 2);
            float b[2];
            cc_cli_bounds(strtof(argv[i + 1], NULL), strtof(argv[i + 2], NULL), b);
            i += 2;
            theta_cut[0] = b[0]; theta_cut[1] = b[1];
        } else if (!strcmp(a, "-p") || !strcmp(a, "--phi")) {
            need(i, 2);
            float b[2];
            cc_cli_bounds(strtof(argv[i + 1], NULL), strtof(argv[i + 2], NULL), b);
            i += 2;
            phi_cut[0] = b[0]; phi_cut[1] = b[1];
        } else if (!strcmp(a, "-b") || !strcmp(a, "--boxLength")) {
            need(i, 1);
            boxLength = strtof(argv[++i], NULL);
            boxLengths.assign(1, boxLength);
            boxLabels.assign(1, argv[i]);
            while (i + 1 < argc && argv[i + 1][0] != '-') { /* extension: further box sizes */
                char* end = NULL;
                const float b = strtof(argv[i + 1], &end);
                if (end == argv[i + 1] || *end != 0) break;
                boxLengths.push_back(b);
                boxLabels.push_back(argv[++i]);
            }
        } else if (!strcmp(a, "-m") || !strcmp(a, "--massDef")) {
            need(i, 1);
            massDef = argv[++i];
        } else if (!strcmp(a, "-h") || !strcmp(a, "--halo")) {
            need(i, 3);
            for (int k = 0; k < 3; ++k) haloPos.push_back(strtof(argv[++i], NULL));
            for (int k = 0; k < 6; ++k) haloProps.push_back(-1.0f); /* properties unknown (main.cpp:286-293) */
        } else if (!strcmp(a, "-f") || !strcmp(a, "--haloFile")) {
            need(i, 1);
            readHaloFile(argv[++i], haloPos, haloTags, haloProps, massDef);
        } else if (!strcmp(a, "-v") || !strcmp(a, "--verbose")) verbose = true;
        else if (!strcmp(a, "--timeit")) timeit = true;
        else if (!strcmp(a, "--overwrite")) overwrite = true;
        else if (!strcmp(a, "--posOnly")) positionOnly = true;
        else if (!strcmp(a, "--forceWriteProps")) forceWriteProps = true;
        else if (!strcmp(a, "--propsOnly")) propsOnly = true;
    }

    /* one output sub-directory per halo of a halo file (main.cpp:320-349); with several box sizes, one tree per size */
    auto halo_dirs_under = [&](const string& root) {
        vector<string> dirs;
        if (customHaloFile) {
            for (size_t h = 0; h < haloTags.size(); ++h) {
                const string sub = root + "halo_" + haloTags[h] + "/";
                DIR* d = opendir(sub.c_str());
                if (d) closedir(d);
                else {
                    mkdir(sub.c_str(), S_IRWXU | S_IRGRP | S_IXGRP | S_IXOTH);
                    if (h % 100 == 0) cout << "Created output directory: " << sub << endl;
                }
                dirs.push_back(sub);
            }
        } else if (customHalo) {
            dirs.push_back(root);
        }
        return dirs;
    };
    const bool separate_passes = getenv("CC_BOX_PASSES") && !strcmp(getenv("CC_BOX_PASSES"), "separate");
    vector<vector<string>> halo_out_dirs_of;
    if (boxLengths.size() <= 1) {
        halo_out_dirs_of.push_back(halo_dirs_under(out_dir));
    } else {
        for (const string& label : boxLabels) {
            const string root = out_dir + "boxLength_" + label + "/";
            mkdir(root.c_str(), S_IRWXU | S_IRGRP | S_IXGRP | S_IXOTH);
            halo_out_dirs_of.push_back(halo_dirs_under(root));
        }
    }

    if (customHalo || customHaloFile) {
        cout << endl << haloPos.size() / 3 << " target halo(s) will be cut out of the lightcone";
        if (verbose) {
            cout << ":" << endl;
            for (size_t k = 0; k < haloTags.size(); ++k) {
                cout << "Halo " << haloTags[k] << ": " << endl;
                for (int i = 0; i < 3; ++i) cout << cart[i] << "=" << haloPos[3 * k + i] << " ";
            }
        }
        if (boxLengths.size() <= 1) cout << endl << "box length: " << boxLength << " arcmin" << endl;
        else {
            cout << endl << "box lengths:";
            for (float b : boxLengths) cout << " " << b;
            cout << (separate_passes ? " arcmin (one pass each; lightcone steps stay resident in HBM between passes)"
                                     : " arcmin (all sizes share one pass over every lightcone step)") << endl;
        }
    } else {
        cout << "theta bounds: " << theta_cut[0] / ARCSEC << " -> " << theta_cut[1] / ARCSEC << " deg" << endl;
        cout << "phi bounds: " << phi_cut[0] / ARCSEC << " -> " << phi_cut[1] / ARCSEC << " deg" << endl;
    }
    cout << "\nverbose is set to " << verbose << endl;
    cout << "timeit is set to " << timeit << endl;
    cout << "overwrite is set to " << overwrite << endl;
    cout << "posOnly is set to " << positionOnly << endl;

    const double start = wall_seconds();
    if (const char* e = getenv("CC_RESIDENT_STEPS")) resident_steps_enable(atoi(e) != 0);
    if (boxLengths.size() > 1 && separate_passes) resident_steps_enable(true);
    if ((customHalo || customHaloFile) && boxLengths.size() > 1 && !separate_passes)
        processLC_boxes(input_lc_dir, halo_out_dirs_of, step_strings, haloPos, haloProps, boxLengths, myrank, numranks,
                        verbose, timeit, overwrite, positionOnly, forceWriteProps, propsOnly);
    else if (customHalo || customHaloFile)
        for (size_t b = 0; b < boxLengths.size(); ++b) {
            if (boxLengths.size() > 1) cout << "\n\n########## box length " << boxLengths[b] << " arcmin ##########" << endl;
            processLC(input_lc_dir, halo_out_dirs_of[b], step_strings, haloPos, haloProps, boxLengths[b], myrank, numranks,
                      verbose, timeit, overwrite, positionOnly, forceWriteProps, propsOnly);
        }
    else
        processLC(input_lc_dir, out_dir, step_strings, theta_cut, phi_cut, myrank, numranks, verbose, timeit, overwrite,
                  positionOnly);
    const double duration = wall_seconds() - start;
    if (timeit) cout << "\nExecution time: " << duration << " s" << endl;
    shutdown_gpus();
    return 0;
}
