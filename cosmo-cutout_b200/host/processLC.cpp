/*
 * host/processLC.cpp -- the two processLC overloads of the reference over the C ABI (see processLC.h).
 *
 * Shape of one step, both use cases:
 *   read     the step's columns into pinned host buffers          (reference: GenericIO, processLC.cpp:102-139 / :861-889)
 *   cutout   GPU r streams x,y,z of its contiguous index shard H2D (double-buffered) through the fused scan
 *            kernel; survivors' other columns are gathered straight from the pinned buffers
 *                                                                  (reference: the loops at :276-310 / :897-909 + :1214-1221 + :1383-1481)
 *   counts   NCCL all-gather of the per-GPU survivor counts + exclusive scan -> each GPU's row offset
 *                                                                  (reference: MPI_Allgather + prefix sum, :322-337 / :1546-1583)
 *   write    one raw little-endian file per column, GPU r's block at byte offset sizeof(T)*offset[r]
 *                                                                  (reference: MPI-IO, :355-427 / :1606-1704)
 * There is no load-balancing redistribution (:974-1125): index shards are equal by construction.  Its one
 * observable side effect -- the silently dropped tail once a rank holds more than 2^24 particles -- is reproduced
 * by default (set CC_KEEP_TAIL=1 to keep every particle), so results equal a single-rank reference run.
 */
#include "processLC.h"

#include <dirent.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "column_reader.h"
#include "cosmo_cutout_b200.h"
#include "util.h"

using std::cout;
using std::endl;
using std::string;
using std::vector;
using namespace cchost;

namespace {

/* ---------------------------------------------------------------- GPUs: one context per "rank" */
struct GpuPool {
    vector<cc_ctx*> ctx;
    ~GpuPool() { for (cc_ctx* c : ctx) cc_destroy(c); }
    void ensure(int numranks) {
        if (numranks < 1 || cc_device_count() < 1) die("lc_cutout: no CUDA device is visible; this build has no CPU path");
        if ((int)ctx.size() == numranks) return;
        for (cc_ctx* c : ctx) cc_destroy(c);
        ctx.clear();
        const int ndev = cc_device_count();
        if (ndev < 1) die("lc_cutout: no CUDA device is visible; this build has no CPU path");
        if (numranks > ndev) die("lc_cutout: " + std::to_string(numranks) + " GPUs requested, " + std::to_string(ndev) + " visible");
        for (int r = 0; r < numranks; ++r) {
            cc_ctx* c = nullptr;
            if (cc_create(r, &c) != CC_OK) die(string("lc_cutout: ") + cc_last_error());
            ctx.push_back(c);
        }
        if (numranks > 1 && cc_comm_init_all(ctx.data(), numranks) != CC_OK) die(string("lc_cutout: ") + cc_last_error());
    }
};
GpuPool g_pool;

void check(int rc) {
    if (rc != CC_OK) die(string("lc_cutout: ") + cc_last_error());
}

/* run f(rank) for every GPU on its own host thread (NCCL collectives of one process must be entered concurrently) */
template <typename F>
void for_each_gpu(int numranks, F f) {
    if (numranks == 1) { f(0); return; }
    vector<std::thread> th;
    for (int r = 0; r < numranks; ++r) th.emplace_back([&, r] { f(r); });
    for (auto& t : th) t.join();
}

/* contiguous index shard of GPU r */
void shard_range(size_t n, int numranks, int r, size_t* lo, size_t* hi) {
    int64_t a, b;
    cc_shard_range((int64_t)n, numranks, r, &a, &b);
    *lo = (size_t)a;
    *hi = (size_t)b;
}

/* ---------------------------------------------------------------- survivors on the host */
struct Survivors {
    vector<float> x, y, z, vx, vy, vz, redshift, theta, phi, theta_u;
    vector<int64_t> id, index;
    vector<int32_t> rotation, replication;
    size_t n = 0;
    void resize(size_t k, bool vel, bool keys) {
        n = k;
        x.resize(k); y.resize(k); z.resize(k); redshift.resize(k); theta.resize(k); phi.resize(k); id.resize(k);
        if (vel) { vx.resize(k); vy.resize(k); vz.resize(k); rotation.resize(k); replication.resize(k); }
        if (keys) { theta_u.resize(k); index.resize(k); }
    }
    /* rows [row, row+k) of src */
    void slice_of(const Survivors& src, size_t row, size_t k, bool vel, bool keys) {
        n = k;
        auto cp = [row, k](auto& dst, const auto& from) { dst.assign(from.begin() + (ptrdiff_t)row, from.begin() + (ptrdiff_t)(row + k)); };
        cp(x, src.x); cp(y, src.y); cp(z, src.z); cp(redshift, src.redshift); cp(theta, src.theta); cp(phi, src.phi); cp(id, src.id);
        if (vel) { cp(vx, src.vx); cp(vy, src.vy); cp(vz, src.vz); cp(rotation, src.rotation); cp(replication, src.replication); }
        if (keys) { cp(theta_u, src.theta_u); cp(index, src.index); }
    }
    cc_cols_out view(bool vel, bool keys) {
        cc_cols_out o;
        memset(&o, 0, sizeof(o));
        o.x = x.data(); o.y = y.data(); o.z = z.data(); o.redshift = redshift.data(); o.theta = theta.data();
        o.phi = phi.data(); o.id = id.data();
        if (vel) { o.vx = vx.data(); o.vy = vy.data(); o.vz = vz.data(); o.rotation = rotation.data(); o.replication = replication.data(); }
        if (keys) { o.theta_u = theta_u.data(); o.index = index.data(); }
        return o;
    }
};

/* ---------------------------------------------------------------- input read-ahead
 * While step i is being cut out and written, a background thread already reads step i+1 into the second set of
 * pinned buffers: once a cutout takes milliseconds the file read is the end-to-end bottleneck.
 * Costs a second set of input buffers in host memory; CC_NO_READAHEAD=1 turns it off. */
class StepReadAhead {
public:
    StepReadAhead() : slot_(0), pending_(false), enabled_(!(getenv("CC_NO_READAHEAD") && string(getenv("CC_NO_READAHEAD")) == "1")) {}
    ~StepReadAhead() { if (pending_) th_.join(); }
    /* the columns of `header`: already being read, or read now */
    StepColumns& take(const string& header, bool positionOnly) {
        if (pending_) {
            th_.join();
            pending_ = false;
            if (pending_header_ == header && pending_pos_ == positionOnly) { slot_ = 1 - slot_; return buf_[slot_]; }
        }
        buf_[slot_].read(header, positionOnly);
        return buf_[slot_];
    }
    /* start reading the step that will be asked for next (into the buffers not handed out last) */
    void start(const string& header, bool positionOnly) {
        if (!enabled_ || pending_) return;
        pending_header_ = header;
        pending_pos_ = positionOnly;
        StepColumns* dst = &buf_[1 - slot_];
        th_ = std::thread([dst, header, positionOnly] { dst->read(header, positionOnly); });
        pending_ = true;
    }

private:
    StepColumns buf_[2];
    int slot_;
    bool pending_, enabled_;
    string pending_header_;
    bool pending_pos_ = false;
    std::thread th_;
};

/* header file of the next step that will be processed after position i (step 499 is skipped), or "" */
string next_header(const string& dir_name, const string& subdirPrefix, const vector<string>& step_strings, size_t i) {
    for (size_t j = i + 1; j < step_strings.size(); ++j) {
        if (atoi(step_strings[j].c_str()) == 499) continue;
        const string step_dir = dir_name + subdirPrefix + step_strings[j];
        /* the same selection as getLCFile (util.cpp:286-292), but silent: anything irregular is left to the step's
         * own turn, where getLCFile reports it exactly as before */
        DIR* dp = opendir(step_dir.c_str());
        if (!dp) return "";
        string file_name;
        int found = 0;
        while (struct dirent* e = readdir(dp)) {
            const string name(e->d_name);
            if (name.find("lc") != string::npos && name.find("#") == string::npos && name.find("SubInput") == string::npos) {
                file_name = name;
                ++found;
            }
        }
        closedir(dp);
        if (found != 1) return "";
        return step_dir + "/" + file_name;
    }
    return "";
}

/* ---------------------------------------------------------------- output files (processLC.cpp:226-264 / :1291-1302) */
struct ColumnFile {
    int fd = -1;
    void open_trunc(const string& path) {
        fd = open(path.c_str(), O_CREAT | O_WRONLY | O_TRUNC, 0644);
        if (fd < 0) die("lc_cutout: cannot create " + path);
    }
    void write_at(const void* buf, size_t bytes, size_t offset) {
        size_t done = 0;
        while (done < bytes) {
            const ssize_t w = pwrite(fd, (const char*)buf + done, bytes - done, (off_t)(offset + done));
            if (w <= 0) die("lc_cutout: write failed");
            done += (size_t)w;
        }
    }
    void close_() { if (fd >= 0) close(fd); fd = -1; }
};

struct CutoutFiles {
    ColumnFile id, x, y, z, vx, vy, vz, redshift, theta, phi, rotation, replication;
    bool vel;
    CutoutFiles(const string& subdir, int step, bool with_vel) : vel(with_vel) {
        auto name = [&](const char* col) { return subdir + "/" + col + "." + std::to_string(step) + ".bin"; };
        id.open_trunc(name("id")); x.open_trunc(name("x")); y.open_trunc(name("y")); z.open_trunc(name("z"));
        theta.open_trunc(name("theta")); phi.open_trunc(name("phi")); redshift.open_trunc(name("redshift"));
        if (vel) {
            vx.open_trunc(name("vx")); vy.open_trunc(name("vy")); vz.open_trunc(name("vz"));
            rotation.open_trunc(name("rotation")); replication.open_trunc(name("replication"));
        }
    }
    /* rows [0,k) of s go to row offset `row` of every file */
    void write_block(const Survivors& s, size_t k, size_t row) {
        id.write_at(s.id.data(), k * 8, row * 8);
        x.write_at(s.x.data(), k * 4, row * 4); y.write_at(s.y.data(), k * 4, row * 4); z.write_at(s.z.data(), k * 4, row * 4);
        theta.write_at(s.theta.data(), k * 4, row * 4); phi.write_at(s.phi.data(), k * 4, row * 4);
        redshift.write_at(s.redshift.data(), k * 4, row * 4);
        if (vel) {
            vx.write_at(s.vx.data(), k * 4, row * 4); vy.write_at(s.vy.data(), k * 4, row * 4); vz.write_at(s.vz.data(), k * 4, row * 4);
            rotation.write_at(s.rotation.data(), k * 4, row * 4); replication.write_at(s.replication.data(), k * 4, row * 4);
        }
    }
    ~CutoutFiles() {
        ColumnFile* all[] = {&id, &x, &y, &z, &vx, &vy, &vz, &redshift, &theta, &phi, &rotation, &replication};
        for (ColumnFile* f : all) f->close_();
    }
};

template <typename T>
void print_vec(const char* label, const vector<T>& v, std::ostream& os = cout) {
    os << label << ": [";
    for (const T& e : v) os << e << ",";
    os << "]" << endl;
}

void print_times(const char* name, const vector<double>& v) {
    cout << name << " = np.array([";
    for (size_t i = 0; i < v.size(); ++i) {
        cout << v[i];
        if (i + 1 < v.size()) cout << ", ";
    }
    cout << "]" << endl;
}

string find_prefix(const string& dir_name, bool print) {
    if (print) cout << "\nReading directory: " << dir_name << endl;
    vector<string> subdirs;
    getLCSubdirs(dir_name, subdirs);
    if (print) {
        cout << "Found subdirs:" << endl;
        for (const string& s : subdirs) cout << s << ' ';
        cout << endl;
    }
    if (subdirs.empty()) die("lc_cutout: no lightcone step sub-directories under " + dir_name);
    const string prefix = subdirPrefixOf(subdirs);
    if (print) cout << "Subdir prefix is: " << prefix << endl;
    return prefix;
}

}  // namespace

/* =====================================================================================================
 * use case 1 (reference processLC.cpp:15-429)
 * ===================================================================================================== */
void processLC(string dir_name, string out_dir, vector<string> step_strings, vector<float> theta_cut,
               vector<float> phi_cut, int myrank, int numranks, bool verbose, bool timeit, bool overwrite,
               bool positionOnly) {
    /* like the reference, use case 1 ignores verbose / timeit / overwrite / positionOnly */
    (void)verbose; (void)timeit; (void)overwrite; (void)positionOnly;
    if (myrank != 0) die("processLC: this implementation is one process driving all GPUs; call it with myrank == 0");
    if (theta_cut.size() != 2 || phi_cut.size() != 2) die("processLC: theta/phi bounds must be {lo, hi}");
    g_pool.ensure(numranks);
    const string subdirPrefix = find_prefix(dir_name, true);

    StepReadAhead reader; /* two sets of pinned buffers, reused across steps */
    for (size_t i = 0; i < step_strings.size(); ++i) {
        const int step = atoi(step_strings[i].c_str());
        if (step == 499) continue; /* z = 0: zero lightcone volume (processLC.cpp:72-73) */
        cout << "\n---------- Working on step " << step_strings[i] << "----------" << endl;
        const string step_dir = dir_name + subdirPrefix + step_strings[i];
        string file_name;
        getLCFile(step_dir, file_name);
        const string header = step_dir + "/" + file_name;
        cout << "Opening file: " << header << endl;
        StepColumns& cols = reader.take(header, false);
        {
            const string nxt = next_header(dir_name, subdirPrefix, step_strings, i);
            if (!nxt.empty()) reader.start(nxt, false);
        }
        const size_t Np = cols.n;
        cout << "Number of elements in lc step at rank " << myrank << ": " << Np << endl;

        /* output sub-directory: must be absent or empty (processLC.cpp:161-184) */
        const string step_subdir = out_dir + subdirPrefix + "Cutout" + step_strings[i];
        {
            vector<string> probe;
            struct stat st;
            if (stat(step_subdir.c_str(), &st) == 0) {
                if (prepStepSubdir(step_subdir, false, false, false) == 2) die("\nDirectory " + step_subdir + " is non-empty");
                cout << "Entered subdir: " << step_subdir << endl;
            } else {
                mkdir(step_subdir.c_str(), S_IRWXU | S_IRGRP | S_IXGRP | S_IXOTH);
                cout << "Created subdir: " << step_subdir << endl;
            }
        }
        CutoutFiles files(step_subdir, step, true);

        cout << "Converting positions..." << endl;
        const float th[2] = {theta_cut[0], theta_cut[1]}, ph[2] = {phi_cut[0], phi_cut[1]};
        for (int r = 0; r < numranks; ++r) { /* enqueue on every GPU, nothing blocks here */
            size_t lo, hi;
            shard_range(Np, numranks, r, &lo, &hi);
            const cc_cols_in v = cols.view(lo);
            check(cc_cutout_const_streamed(g_pool.ctx[r], (int64_t)(hi - lo), &v, (int64_t)lo, th, ph));
        }
        vector<int64_t> all_counts((size_t)numranks), offsets((size_t)numranks);
        for_each_gpu(numranks, [&](int r) {
            vector<int64_t> mine((size_t)numranks);
            int64_t off = 0;
            check(cc_exchange_counts(g_pool.ctx[r], mine.data(), nullptr, &off));
            offsets[(size_t)r] = off;
            if (r == 0) all_counts = mine;
        });
        print_vec("rank object counts", all_counts);
        print_vec("rank offsets", offsets);

        for_each_gpu(numranks, [&](int r) {
            Survivors s;
            const size_t k = (size_t)all_counts[(size_t)r];
            s.resize(k, true, false);
            const cc_cols_out o = s.view(true, false);
            if (k) check(cc_cutout_fetch(g_pool.ctx[r], 0, 0, (int64_t)k, &o));
            files.write_block(s, k, (size_t)offsets[(size_t)r]);
        });
    }
}

/* =====================================================================================================
 * use case 2 (reference processLC.cpp:440-1757)
 * ===================================================================================================== */
void processLC(string dir_name, vector<string> out_dirs, vector<string> step_strings, vector<float> halo_pos,
               vector<float> halo_props, float boxLength, int myrank, int numranks, bool verbose, bool timeit,
               bool overwrite, bool positionOnly, bool forceWriteProps, bool propsOnly) {
    if (myrank != 0) die("processLC: this implementation is one process driving all GPUs; call it with myrank == 0");
    const string subdirPrefix = find_prefix(dir_name, true);
    const int numHalos = (int)(halo_pos.size() / 3);
    if ((int)out_dirs.size() != numHalos) die("processLC: one output directory per halo is required");

    /* ---------------- per-halo bounds + rotation (host, O(1) per halo; reference :512-770) ---------------- */
    cout << "\n\n---------- Setting up for coordinate rotation ----------" << endl;
    vector<cc_halo> halos((size_t)numHalos);
    const int numProps = numHalos ? (int)(halo_props.size() / (size_t)numHalos) : 0;
    if (numHalos && numProps != 3 && numProps != 6) {
        cout << "Something went wrong... " << numProps << " properties found per halo in input vectors. "
             << "Is massDef correct? Please contact developer." << endl;
        exit(EXIT_FAILURE);
    }
    for (int h = 0; h < numHalos; ++h) {
        const bool printHalo = (numHalos < 20) || (h % 100 == 0);
        const float pos[3] = {halo_pos[3 * h], halo_pos[3 * h + 1], halo_pos[3 * h + 2]};
        cc_halo_info info;
        check(cc_halo_setup(pos, boxLength, &info));
        halos[(size_t)h] = info.halo;
        const cc_halo& H = info.halo;
        if (printHalo) {
            const float dtheta = info.half_box_rad, dphi = dtheta;
            cout << "\n--- Target halo " << h << " ---" << endl;
            cout << "theta bounds set to: " << H.theta_cut[0] / ARCSEC << "deg -> " << H.theta_cut[1] / ARCSEC << "deg" << endl;
            cout << "phi bounds set to: " << H.phi_cut[0] / ARCSEC << "deg -> " << H.phi_cut[1] / ARCSEC << "deg" << endl;
            cout << "theta-phi bounds result in box width of " << tanf(dtheta) * info.halo_r * 2
                 << " Mpc/h at distance to halo of " << info.halo_r << " Mpc/h" << endl
                 << "        = " << dtheta * 2 * 180.0 / PI << "deg x " << dphi * 2 * 180.0 / PI << "deg field of view" << endl;
            cout << "\nFinding axis of rotation to move (" << pos[0] << ", " << pos[1] << ", " << pos[2] << ") to ("
                 << info.halo_r << ", " << 0 << ", " << 0 << ")" << endl;
            cout << "Rotation is " << info.beta * (180 / PI) << "deg about axis k = (" << info.k[0] << ", " << info.k[1]
                 << ", " << info.k[2] << ")" << endl;
            if (verbose) {
                const float* R = info.R;
                const float* Ri = info.R_inv;
                cout << "\nRotation Matrix is " << endl
                     << "{ " << R[0] << ", " << R[1] << ", " << R[2] << "}" << endl
                     << "{ " << R[3] << ", " << R[4] << ", " << R[5] << "}" << endl
                     << "{ " << R[6] << ", " << R[7] << ", " << R[8] << "}" << endl;
                cout << "Inverted Rotation Matrix is " << endl
                     << "{ " << Ri[0] << ", " << Ri[1] << ", " << Ri[2] << "}" << endl
                     << "{ " << Ri[3] << ", " << Ri[4] << ", " << Ri[5] << "}" << endl
                     << "{ " << Ri[6] << ", " << Ri[7] << ", " << Ri[8] << "}" << endl;
                const char* nm = "ABCD";
                cout << "\nRotated cartesian vectors pointing toward FOV corners are" << endl;
                for (int c = 0; c < 4; ++c)
                    cout << nm[c] << " = {" << info.corners_rot[3 * c] << ", " << info.corners_rot[3 * c + 1] << ", "
                         << info.corners_rot[3 * c + 2] << "}" << endl;
                cout << "\nAngular positions of the FOV corners in degrees are" << endl;
                for (int c = 0; c < 4; ++c)
                    cout << nm[c] << " = {" << info.corners_sph[2 * c] / ARCSEC << ", " << info.corners_sph[2 * c + 1] / ARCSEC << "}" << endl;
            }
            cout << "\nrough theta bounds set to: " << H.theta_rough[0] / ARCSEC << "deg -> " << H.theta_rough[1] / ARCSEC << "deg" << endl;
            cout << "rough phi bounds set to: " << H.phi_rough[0] / ARCSEC << "deg -> " << H.phi_rough[1] / ARCSEC << "deg" << endl;
        }
        /* properties.csv (reference :740-769) */
        const string props_name = out_dirs[(size_t)h] + "/properties.csv";
        if (!does_file_exist(props_name) || forceWriteProps) {
            std::ofstream pf(props_name.c_str());
            pf << "#halo_redshift" << ", " << "halo_lc_shell" << ", ";
            if (numProps == 6) pf << "sod_halo_mass" << ", " << "sod_halo_radius" << ", " << "sod_halo_cdelta" << ", " << "sod_halo_cdelta_error";
            else pf << "fof_halo_mass";
            pf << ", " << "halo_lc_x" << ", " << "halo_lc_y" << ", " << "halo_lc_z" << ", " << "boxRadius_Mpc" << ", " << "boxRadius_arcsec" << "\n";
            for (int i = 0; i < numProps; ++i) pf << halo_props[(size_t)(numProps * h + i)] << ", ";
            for (int i = 0; i < 3; ++i) pf << pos[i] << ", ";
            pf << tanf(info.half_box_rad) * info.halo_r << ", " << info.half_box_rad * 180.0 / PI * ARCSEC << "\n";
            pf.close();
            if (printHalo) cout << "wrote halo info to properties.csv" << endl;
        } else if (printHalo) {
            cout << "properties.csv already exists; skipping write" << endl;
        }
    }
    if (propsOnly) return;

    g_pool.ensure(numranks);
    const bool keep_tail = getenv("CC_KEEP_TAIL") && string(getenv("CC_KEEP_TAIL")) == "1";
    vector<double> read_times, redist_times, sort_times, cutout_times, write_times;
    StepReadAhead reader;

    for (size_t i = 0; i < step_strings.size(); ++i) {
        double start = wall_seconds();
        const int step = atoi(step_strings[i].c_str());
        if (step == 499) continue; /* reference :805-806 */
        cout << "\n=================================================" << endl;
        cout << "============== Working on step " << step_strings[i] << " ==============\n" << endl;
        const string step_dir = dir_name + subdirPrefix + step_strings[i];
        string file_name;
        cout << "getting lc file..." << endl;
        getLCFile(step_dir, file_name);
        const string header = step_dir + "/" + file_name;
        cout << "Opening file: " << header << endl;
        StepColumns& cols = reader.take(header, positionOnly);
        {
            const string nxt = next_header(dir_name, subdirPrefix, step_strings, i);
            if (!nxt.empty()) reader.start(nxt, positionOnly);
        }
        const size_t Np = cols.n;
        double duration = wall_seconds() - start;
        if (timeit) cout << "Read time: " << duration << " s" << endl;
        read_times.push_back(duration);
        redist_times.push_back(0.0); /* no redistribution: index shards are equal by construction */
        sort_times.push_back(0.0);   /* no full sort: only survivors are ordered, inside the cutout */

        /* the reference's load balancer keeps only a prefix of each rank's block once Np > 2^24 */
        const int64_t n_limit = keep_tail ? -1 : cc_ref_scatter_kept((int64_t)Np, 1);
        if (n_limit >= 0 && (size_t)n_limit != Np)
            cout << "note: reproducing the reference's redistribution tail drop: last " << (Np - (size_t)n_limit)
                 << " of " << Np << " particles are not considered (CC_KEEP_TAIL=1 keeps them)" << endl;

        /* output directories first: a halo whose directory is non-empty (and no --overwrite) is skipped (:1261-1274) */
        vector<int> active;
        vector<string> subdirs_of((size_t)numHalos);
        for (int h = 0; h < numHalos; ++h) {
            const bool printHalo = (numHalos < 20) || (h % 100 == 0);
            subdirs_of[(size_t)h] = out_dirs[(size_t)h] + subdirPrefix + "Cutout" + step_strings[i];
            const int error = prepStepSubdir(subdirs_of[(size_t)h], overwrite, printHalo, verbose);
            if (error == 1) exit(EXIT_FAILURE);
            if (error == 2) continue;
            active.push_back(h);
        }
        if (active.empty()) continue;

        /* one pass over the step for all active halos, every GPU on its own shard */
        start = wall_seconds();
        vector<cc_halo> act_halos;
        for (int h : active) act_halos.push_back(halos[(size_t)h]);
        const int A = (int)act_halos.size();
        for (int r = 0; r < numranks; ++r) {
            size_t lo, hi;
            shard_range(Np, numranks, r, &lo, &hi);
            const cc_cols_in v = cols.view(lo);
            check(cc_cutout_halo_streamed(g_pool.ctx[r], (int64_t)(hi - lo), &v, (int64_t)lo, A, act_halos.data(),
                                          positionOnly ? 1 : 0, n_limit));
        }
        vector<int64_t> all_counts((size_t)numranks * (size_t)A), all_rough((size_t)numranks * (size_t)A);
        for_each_gpu(numranks, [&](int r) {
            vector<int64_t> c((size_t)numranks * (size_t)A), g((size_t)numranks * (size_t)A), off((size_t)A);
            check(cc_exchange_counts(g_pool.ctx[r], c.data(), g.data(), off.data()));
            if (r == 0) { all_counts = c; all_rough = g; }
        });
        duration = wall_seconds() - start;
        if (timeit) cout << "cutout computation time (all halos of the step): " << duration << " s" << endl;

        /* the survivors of ALL halos of the step, one transfer per column and GPU (a --haloFile run has hundreds of
         * halos per step); halo ai of GPU r sits at rows [row_of[r][ai], +count) */
        vector<Survivors> fetched((size_t)numranks);
        vector<vector<size_t>> row_of((size_t)numranks, vector<size_t>((size_t)A + 1, 0));
        for (int r = 0; r < numranks; ++r)
            for (int ai = 0; ai < A; ++ai)
                row_of[(size_t)r][(size_t)ai + 1] = row_of[(size_t)r][(size_t)ai] + (size_t)all_counts[(size_t)r * (size_t)A + (size_t)ai];
        for_each_gpu(numranks, [&](int r) {
            const size_t k = row_of[(size_t)r][(size_t)A];
            fetched[(size_t)r].resize(k, !positionOnly, true);
            const cc_cols_out o = fetched[(size_t)r].view(!positionOnly, true);
            check(cc_cutout_fetch_all(g_pool.ctx[r], (int64_t)k, &o));
        });

        /* per halo: merge the GPUs' lists and write the column files.  A --haloFile step has hundreds of halos and
         * twelve small files each: halos are written by a few threads at a time, each logging into its own buffer,
         * and the buffers are printed in halo order */
        auto write_halo = [&](int ai, std::ostream& log) -> double {
            const int h = active[(size_t)ai];
            const bool printHalo = (numHalos < 20) || (h % 100 == 0);
            if (printHalo) log << "\n---------- cutout at halo " << h << "----------" << endl;
            const double t0 = wall_seconds();
            vector<int64_t> np_count((size_t)numranks), np_rough((size_t)numranks), np_offset((size_t)numranks);
            int64_t tot = 0, tot_rough = 0;
            for (int r = 0; r < numranks; ++r) {
                np_count[(size_t)r] = all_counts[(size_t)r * (size_t)A + (size_t)ai];
                np_rough[(size_t)r] = all_rough[(size_t)r * (size_t)A + (size_t)ai];
                np_offset[(size_t)r] = tot;
                tot += np_count[(size_t)r];
                tot_rough += np_rough[(size_t)r];
            }
            if (printHalo) {
                log << "(" << tot_rough << ") " << tot << " total particles in (rough) cutout" << endl;
                print_vec("rank object rough counts", np_rough, log);
                print_vec("rank object counts", np_count, log);
                print_vec("rank offsets", np_offset, log);
                log << "Writing files..." << endl;
            }
            /* every GPU's survivors arrive ordered by (theta_u, index); a single-rank reference run orders the whole
             * step that way (:1214-1221), so merge the per-GPU lists on those keys */
            vector<Survivors> parts((size_t)numranks);
            for (int r = 0; r < numranks; ++r)
                parts[(size_t)r].slice_of(fetched[(size_t)r], row_of[(size_t)r][(size_t)ai], (size_t)np_count[(size_t)r], !positionOnly, true);
            CutoutFiles files(subdirs_of[(size_t)h], step, !positionOnly);
            if (numranks == 1) {
                files.write_block(parts[0], parts[0].n, 0);
            } else {
                vector<int32_t> src_rank((size_t)tot);
                vector<int64_t> src_row((size_t)tot);
                vector<const float*> keys_t((size_t)numranks);
                vector<const int64_t*> keys_i((size_t)numranks);
                for (int r = 0; r < numranks; ++r) { keys_t[(size_t)r] = parts[(size_t)r].theta_u.data(); keys_i[(size_t)r] = parts[(size_t)r].index.data(); }
                check(cc_merge_order(numranks, np_count.data(), keys_t.data(), keys_i.data(), src_rank.data(), src_row.data()));
                Survivors m;
                m.resize((size_t)tot, !positionOnly, false);
                for (size_t o = 0; o < (size_t)tot; ++o) {
                    const Survivors& s = parts[(size_t)src_rank[o]];
                    const size_t j = (size_t)src_row[o];
                    m.x[o] = s.x[j]; m.y[o] = s.y[j]; m.z[o] = s.z[j]; m.redshift[o] = s.redshift[j];
                    m.theta[o] = s.theta[j]; m.phi[o] = s.phi[j]; m.id[o] = s.id[j];
                    if (!positionOnly) {
                        m.vx[o] = s.vx[j]; m.vy[o] = s.vy[j]; m.vz[o] = s.vz[j];
                        m.rotation[o] = s.rotation[j]; m.replication[o] = s.replication[j];
                    }
                }
                files.write_block(m, (size_t)tot, 0);
            }
            const double wdur = wall_seconds() - t0;
            if (timeit && printHalo) log << "write time: " << wdur << " s" << endl;
            return wdur;
        };
        const int writers = std::max(1, std::min({A, 16, (int)std::thread::hardware_concurrency()}));
        for (int a0 = 0; a0 < A; a0 += writers) {
            const int a1 = std::min(A, a0 + writers);
            vector<std::ostringstream> logs((size_t)(a1 - a0));
            vector<double> wd((size_t)(a1 - a0), 0.0);
            if (a1 - a0 == 1) {
                wd[0] = write_halo(a0, logs[0]);
            } else {
                vector<std::thread> th;
                for (int ai = a0; ai < a1; ++ai)
                    th.emplace_back([&, ai] { wd[(size_t)(ai - a0)] = write_halo(ai, logs[(size_t)(ai - a0)]); });
                for (auto& t : th) t.join();
            }
            for (int ai = a0; ai < a1; ++ai) {
                cout << logs[(size_t)(ai - a0)].str();
                cutout_times.push_back(duration / A);
                write_times.push_back(wd[(size_t)(ai - a0)]);
            }
        }
        cout.flush();
    }
    if (timeit) { /* reference :1718-1756: arrays ready to paste into python */
        cout << endl;
        print_times("read_times", read_times);
        print_times("redist_times", redist_times);
        print_times("sort_times", sort_times);
        print_times("cutout_times", cutout_times);
        print_times("write_times", write_times);
    }
}
