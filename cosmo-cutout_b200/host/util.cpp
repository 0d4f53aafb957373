/* host/util.cpp -- see util.h.  Citations are to the reference's src/util.cpp. */
#include "util.h"

#include <dirent.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <thread>

namespace cchost {

using std::cout;
using std::endl;
using std::string;
using std::vector;

namespace {
const std::thread::id g_main_thread = std::this_thread::get_id();
}

/* Fatal: message to stdout, non-zero exit (where the reference calls MPI_Abort / exit(EXIT_FAILURE)).  From a worker
 * thread -- one host thread per GPU around the count exchange, the per-halo writers, the read-ahead -- exit() would run
 * the atexit handlers (the CUDA runtime's teardown among them) while the other threads are still inside CUDA calls or a
 * collective: leave without them. */
void die(const string& msg) {
    cout << msg << endl;
    cout.flush();
    fflush(stdout);
    if (std::this_thread::get_id() != g_main_thread) quick_exit(EXIT_FAILURE);
    exit(EXIT_FAILURE);
}

double wall_seconds() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

bool does_file_exist(const string& filename) {
    std::ifstream f(filename.c_str());
    return f.good();
}

void readHaloFile(const string& haloFileName, vector<float>& haloPos, vector<string>& haloTags,
                  vector<float>& haloProps, const string& massDef) {
    /* util.cpp:172-175: the file is a flat stream of whitespace-separated tokens */
    std::ifstream in(haloFileName.c_str());
    vector<string> tok((std::istream_iterator<string>(in)), std::istream_iterator<string>());
    size_t rowLen;
    if (massDef == "fof") rowLen = 7;
    else if (massDef == "sod") rowLen = 10;
    else die("Unknown mass definition " + massDef + ". Valid options are 'fof' or 'sod'");
    if (tok.size() % rowLen != 0) { /* util.cpp:186-197 */
        if (massDef == "sod")
            die("\nWhen massDef=='sod', each halo position given in input file must have 10 quantities in the "
                "space-delimited form: id redshift shell m200c r200c cdelta cdelta_err x y z ");
        die("\nWhen massDef=='fof', each halo position given in input file must have 7 quantities in the "
            "space-delimited form: id redshift shell m_fof x y z ");
    }
    const size_t nprops = rowLen - 4; /* everything between the tag and x,y,z: 6 (sod) or 3 (fof) */
    for (size_t i = 0; i < tok.size() / rowLen; ++i) {
        const size_t b = i * rowLen;
        haloTags.push_back(tok[b]);
        for (size_t k = 1; k <= nprops; ++k) haloProps.push_back(strtof(tok[b + k].c_str(), 0));
        for (size_t k = rowLen - 3; k < rowLen; ++k) haloPos.push_back(strtof(tok[b + k].c_str(), 0));
    }
}

int getLCSubdirs(const string& dir, vector<string>& subdirs) {
    DIR* dp = opendir(dir.c_str());
    if (!dp) {
        cout << "Error(" << errno << ") opening lightcone data subdirs at " << dir << endl;
        return errno;
    }
    while (struct dirent* e = readdir(dp)) {
        const string name(e->d_name);
        if (name.find("lc") != string::npos || name.find("STEP") != string::npos) subdirs.push_back(name);
    }
    closedir(dp);
    return 0;
}

int getLCFile(const string& dir, string& file) {
    DIR* dp = opendir(dir.c_str());
    if (!dp) {
        cout << "Error(" << errno << ") opening lightcone data files at " << dir << endl;
        return errno;
    }
    vector<string> files; /* util.cpp:286-292: contains "lc", no '#', not a .SubInput file */
    while (struct dirent* e = readdir(dp)) {
        const string name(e->d_name);
        if (name.find("lc") != string::npos && name.find("#") == string::npos && name.find("SubInput") == string::npos)
            files.push_back(name);
    }
    closedir(dp);
    if (files.empty()) die("\nNo valid header files found in dir" + dir);
    if (files.size() > 1)
        die("Too many header files in directory " + dir +
            ". LC Output files should be separated by step-respective subdirectories");
    file = files[0];
    return 0;
}

int getLCSteps(int maxStep, int minStep, const string& dir, vector<string>& step_strings) {
    vector<string> subdirs;
    getLCSubdirs(dir, subdirs);
    vector<int> avail; /* util.cpp:349-358: the number starting at the first digit of each name */
    for (const string& s : subdirs) {
        const size_t j = s.find_first_of("0123456789");
        if (j != string::npos) avail.push_back(atoi(s.substr(j).c_str()));
    }
    std::sort(avail.begin(), avail.end());
    for (int st : avail)
        if (st <= maxStep && st >= minStep) step_strings.push_back(std::to_string(st));
    return 0;
}

string subdirPrefixOf(const vector<string>& subdirs) {
    if (subdirs.empty()) return string();
    const size_t j = subdirs[0].find_first_of("0123456789");
    return j == string::npos ? string() : subdirs[0].substr(0, j);
}

int prepStepSubdir(const string& subdir_name, bool overwrite, bool print, bool verbose) {
    DIR* dir = opendir(subdir_name.c_str());
    if (!dir) { /* util.cpp:409-414 */
        mkdir(subdir_name.c_str(), S_IRWXU | S_IRGRP | S_IXGRP | S_IXOTH);
        if (print) cout << "Created subdir: " << subdir_name << endl;
        return 0;
    }
    vector<string> entries;
    while (struct dirent* e = readdir(dir)) {
        if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
        entries.push_back(e->d_name);
    }
    closedir(dir);
    if (entries.empty()) { /* util.cpp:424-429 */
        if (print) cout << "Entered subdir: " << subdir_name << endl;
        return 0;
    }
    if (!overwrite) { /* util.cpp:432-438 */
        if (print) cout << "\nDirectory " << subdir_name << " is non-empty; skipping halo" << endl;
        return 2;
    }
    if (print) { /* util.cpp:441-475 */
        cout << "\nDirectory " << subdir_name << " is non-empty";
        cout << " and --overwrite flag was passed; removing binary files here" << endl;
    }
    for (const string& name : entries) {
        const size_t dot = name.rfind('.');
        if (dot == string::npos || name.substr(dot) != ".bin") {
            if (print)
                cout << "\nDirectory contains non-cutout files! Not deleting anything. Maybe passed wrong path?" << endl;
            return 1;
        }
    }
    for (const string& name : entries) {
        const string full = subdir_name + "/" + name;
        if (verbose && print) cout << "\nDELETING " << full << endl;
        remove(full.c_str());
    }
    return 0;
}

float aToZ(float a) { return (1.0f / a) - 1.0f; }

int zToStep(float z, int totSteps, float maxZ) {
    /* util.cpp:520-526: all-float arithmetic, floor toward a = 0 */
    const float amin = 1 / (maxZ + 1);
    const float amax = 1.0;
    const float adiff = (amax - amin) / (totSteps);
    const float a = 1 / (1 + z);
    return (int)floorf((a - amin) / adiff);
}

float stepToZ(float step, int totSteps, float maxZ) {
    const float amin = 1 / (maxZ + 1);
    const float amax = 1.0;
    const float adiff = (amax - amin) / (totSteps);
    const float a = amin + (adiff * step);
    return 1 / a - 1;
}

}  // namespace cchost
