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
    /* no destructor: a namespace-scope object would call cc_destroy (stream syncs, ncclCommDestroy, cudaFree) from a
     * static destructor, after the CUDA runtime's own atexit teardown or -- after die() on a worker thread -- next to
     * threads still inside a collective.  cchost::shutdown_gpus() tears the pool down explicitly; otherwise the
     * process exit reclaims everything. */
    void destroy() { for (cc_ctx* c : ctx) cc_destroy(c); ctx.clear(); }
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
GpuPool& g_pool = *new GpuPool();

void check_at(int rc, const char* what) {
    if (rc != CC_OK) die(string("lc_cutout: ") + cc_last_error() + " [" + what + " returned " + std::to_string(rc) + "]");
}
#define check(call) check_at((call), #call)

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
        /* header parse + pinned (re)allocation happen HERE, on the calling thread and before any GPU work of the
         * current step is enqueued: cudaHostAlloc / cudaFreeHost synchronise every context and must not run next to
         * a kernel that waits on another GPU (the count exchange); the background thread only reads the file */
        dst->prepare(header, positionOnly);
        th_ = std::thread([dst] { dst->load(); });
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

/* ---------------------------------------------------------------- resident lightcone steps across processLC calls
 * processLC cuts with ONE boxLength per call (processLC.h:45-48), so a job with several box sizes -- BASELINE config 5,
 * `lc_cutout --haloFile f --boxLength 5 10 20 30` -- calls it once per size over the same lightcone steps.  With the
 * cache enabled the first call uploads every step it reads to HBM (all ten columns of each GPU's index shard, one
 * resident-step slot per lightcone step, cc_step_select) and later calls find them there: no file read, no
 * host -> device copy, the cutout runs on the resident columns.  Entries are keyed by header path, size and mtime (of
 * the header and of the x column of a flat store), and replaced least-recently-used first when HBM runs short
 * (44 B + a possible 32 B payload per particle and GPU).  A step that does not fit is streamed as before. */
struct ResidentSteps {
    struct Entry { string key; int slot; size_t n; bool pos_only; uint64_t last_use; };
    bool enabled = false;
    vector<Entry> entries;
    vector<int> free_slots;
    int next_slot = 1; /* slot 0 is the streaming slot */
    uint64_t clock = 0;
    int64_t hits = 0, misses = 0, not_fitting = 0;
};
ResidentSteps& g_resident = *new ResidentSteps();

string resident_key(const string& header, int numranks) {
    std::ostringstream k;
    struct stat st;
    k << header << "|" << numranks;
    const string files[2] = {header, header + "#x"};
    for (const string& f : files)
        if (stat(f.c_str(), &st) == 0) k << "|" << (long long)st.st_size << "." << (long long)st.st_mtim.tv_sec << "." << st.st_mtim.tv_nsec;
    return k.str();
}

void select_slot_everywhere(int numranks, int slot) {
    for (int r = 0; r < numranks; ++r) check(cc_step_select(g_pool.ctx[r], slot));
}

/* the step is resident: make it every GPU's current step */
bool resident_lookup(const string& header, bool positionOnly, int numranks, size_t* Np) {
    if (!g_resident.enabled) return false;
    const string key = resident_key(header, numranks);
    for (auto& e : g_resident.entries)
        if (e.key == key && (!e.pos_only || positionOnly)) {
            e.last_use = ++g_resident.clock;
            select_slot_everywhere(numranks, e.slot);
            *Np = e.n;
            ++g_resident.hits;
            return true;
        }
    return false;
}

/* upload a step that was just read; false (slot 0 selected) when it does not fit */
bool resident_insert(const string& header, bool positionOnly, int numranks, const StepColumns& cols) {
    if (!g_resident.enabled) return false;
    ++g_resident.misses;
    const size_t shard = (cols.n + (size_t)numranks - 1) / (size_t)numranks;
    const int64_t need = (int64_t)shard * ((positionOnly ? 24 : 44) + 32) + ((int64_t)3 << 30);
    auto fits = [&]() {
        for (int r = 0; r < numranks; ++r) {
            int64_t free_b = 0;
            check(cc_device_mem(g_pool.ctx[r], &free_b, nullptr));
            if (free_b < need) return false;
        }
        return true;
    };
    while (!fits() && !g_resident.entries.empty()) { /* replace the least recently used step */
        size_t victim = 0;
        for (size_t i = 1; i < g_resident.entries.size(); ++i)
            if (g_resident.entries[i].last_use < g_resident.entries[victim].last_use) victim = i;
        const int slot = g_resident.entries[victim].slot;
        for (int r = 0; r < numranks; ++r) { check(cc_step_select(g_pool.ctx[r], slot)); check(cc_step_release(g_pool.ctx[r])); }
        g_resident.free_slots.push_back(slot);
        g_resident.entries.erase(g_resident.entries.begin() + (ptrdiff_t)victim);
    }
    if (!fits()) {
        ++g_resident.not_fitting;
        select_slot_everywhere(numranks, 0);
        return false;
    }
    int slot;
    if (!g_resident.free_slots.empty()) { slot = g_resident.free_slots.back(); g_resident.free_slots.pop_back(); }
    else slot = g_resident.next_slot++;
    for_each_gpu(numranks, [&](int r) {
        size_t lo, hi;
        shard_range(cols.n, numranks, r, &lo, &hi);
        const cc_cols_in v = cols.view(lo);
        check(cc_step_select(g_pool.ctx[r], slot));
        check(cc_step_upload(g_pool.ctx[r], (int64_t)(hi - lo), &v, (int64_t)lo));
    });
    g_resident.entries.push_back(ResidentSteps::Entry{resident_key(header, numranks), slot, cols.n, positionOnly, ++g_resident.clock});
    return true;
}

bool resident_has(const string& header, bool positionOnly, int numranks) {
    if (!g_resident.enabled) return false;
    const string key = resident_key(header, numranks);
    for (const auto& e : g_resident.entries)
        if (e.key == key && (!e.pos_only || positionOnly)) return true;
    return false;
}

int64_t uploaded_bytes(int numranks) {
    int64_t b = 0;
    for (int r = 0; r < numranks && r < (int)g_pool.ctx.size(); ++r) b += cc_bytes_uploaded(g_pool.ctx[r]);
    return b;
}

void report_transfers(int numranks, int64_t bytes0, int64_t hits0, int64_t misses0) {
    cout << "\nstep columns copied host->device in this call: " << (uploaded_bytes(numranks) - bytes0) << " bytes";
    if (g_resident.enabled)
        cout << " (resident steps: " << (g_resident.hits - hits0) << " found in HBM, " << (g_resident.misses - misses0)
             << " read and uploaded, " << g_resident.entries.size() << " resident now)";
    cout << endl;
}

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

namespace cchost {
void resident_steps_enable(bool on) { g_resident.enabled = on; }
void shutdown_gpus() {
    g_resident.entries.clear();
    g_pool.destroy();
}
}  // namespace cchost

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
    const int64_t bytes0 = uploaded_bytes(numranks), hits0 = g_resident.hits, misses0 = g_resident.misses;
    for (size_t i = 0; i < step_strings.size(); ++i) {
        const int step = atoi(step_strings[i].c_str());
        if (step == 499) continue; /* z = 0: zero lightcone volume (processLC.cpp:72-73) */
        cout << "\n---------- Working on step " << step_strings[i] << "----------" << endl;
        const string step_dir = dir_name + subdirPrefix + step_strings[i];
        string file_name;
        getLCFile(step_dir, file_name);
        const string header = step_dir + "/" + file_name;
        cout << "Opening file: " << header << endl;
        size_t Np = 0;
        StepColumns* cols = nullptr;
        bool resident = resident_lookup(header, false, numranks, &Np);
        if (!resident) {
            cols = &reader.take(header, false);
            Np = cols->n;
            resident = resident_insert(header, false, numranks, *cols);
            if (!resident && g_resident.enabled) select_slot_everywhere(numranks, 0);
        }
        {
            const string nxt = next_header(dir_name, subdirPrefix, step_strings, i);
            if (!nxt.empty() && !resident_has(nxt, false, numranks)) reader.start(nxt, false);
        }
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
            if (resident) {
                check(cc_cutout_const(g_pool.ctx[r], th, ph));
            } else {
                const cc_cols_in v = cols->view(lo);
                check(cc_cutout_const_streamed(g_pool.ctx[r], (int64_t)(hi - lo), &v, (int64_t)lo, th, ph));
            }
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
    report_transfers(numranks, bytes0, hits0, misses0);
}

/* =====================================================================================================
 * use case 2 (reference processLC.cpp:440-1757)
 * ===================================================================================================== */
/* the body of use case 2 with ONE BOX LENGTH PER HALO: the reference's entry point passes the same value for every
 * halo; cchost::processLC_boxes passes the (halo, box length) pairs of several calls so that they share one device
 * pass per step (every pair is an independent cutout with its own bounds, rotation and output directory) */
static void process_halo_list(const string& dir_name, const vector<string>& out_dirs, const vector<string>& step_strings,
                              const vector<float>& halo_pos, const vector<float>& halo_props,
                              const vector<float>& box_of_halo, int myrank, int numranks, bool verbose, bool timeit,
                              bool overwrite, bool positionOnly, bool forceWriteProps, bool propsOnly) {
    if (myrank != 0) die("processLC: this implementation is one process driving all GPUs; call it with myrank == 0");
    const string subdirPrefix = find_prefix(dir_name, true);
    const int numHalos = (int)(halo_pos.size() / 3);
    if ((int)out_dirs.size() != numHalos) die("processLC: one output directory per halo is required");
    if ((int)box_of_halo.size() != numHalos) die("processLC: one box length per halo is required");

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
        check(cc_halo_setup(pos, box_of_halo[(size_t)h], &info));
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
    const int64_t bytes0 = uploaded_bytes(numranks), hits0 = g_resident.hits, misses0 = g_resident.misses;
    int64_t nan_theta_total = 0;

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
        size_t Np = 0;
        StepColumns* cols = nullptr;
        bool resident = resident_lookup(header, positionOnly, numranks, &Np);
        if (!resident) {
            cols = &reader.take(header, positionOnly);
            Np = cols->n;
            resident = resident_insert(header, positionOnly, numranks, *cols);
            if (!resident && g_resident.enabled) select_slot_everywhere(numranks, 0);
        }
        {
            const string nxt = next_header(dir_name, subdirPrefix, step_strings, i);
            if (!nxt.empty() && !resident_has(nxt, positionOnly, numranks)) reader.start(nxt, positionOnly);
        }
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
            if (resident) {
                check(cc_cutout_halo(g_pool.ctx[r], A, act_halos.data(), positionOnly ? 1 : 0, n_limit));
            } else {
                const cc_cols_in v = cols->view(lo);
                check(cc_cutout_halo_streamed(g_pool.ctx[r], (int64_t)(hi - lo), &v, (int64_t)lo, A, act_halos.data(),
                                              positionOnly ? 1 : 0, n_limit));
            }
        }
        vector<int64_t> all_counts((size_t)numranks * (size_t)A), all_rough((size_t)numranks * (size_t)A);
        for_each_gpu(numranks, [&](int r) {
            vector<int64_t> c((size_t)numranks * (size_t)A), g((size_t)numranks * (size_t)A), off((size_t)A);
            check(cc_exchange_counts(g_pool.ctx[r], c.data(), g.data(), off.data()));
            if (r == 0) { all_counts = c; all_rough = g; }
        });
        {
            /* positions whose un-rotated theta is NaN (NaN coordinate, z = +-Inf, the origin): the reference orders all
             * particles by that angle with std::stable_sort / std::lower_bound (processLC.cpp:1214-1221, :1334-1337),
             * which a NaN key breaks -- what it then writes for the OTHER particles is unspecified, so say so */
            int64_t nan_step = 0;
            for (int r = 0; r < numranks; ++r) {
                int64_t st[CC_STATS_EX_FIELDS] = {};
                check(cc_last_stats_ex(g_pool.ctx[r], st, CC_STATS_EX_FIELDS));
                nan_step += st[4];
            }
            if (nan_step) {
                nan_theta_total += nan_step;
                cout << "WARNING: step " << step_strings[i] << " holds " << nan_step << " particle(s) whose theta is NaN (NaN or "
                     << "infinite coordinates, or a particle at the origin).  They are in no cutout here; the reference sorts "
                     << "by theta, and its output for such a step is unspecified -- results may differ from it." << endl;
            }
        }
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
    report_transfers(numranks, bytes0, hits0, misses0);
    if (nan_theta_total) cout << "WARNING: " << nan_theta_total << " particle(s) with NaN theta in this run (see above)" << endl;
    if (timeit) { /* reference :1718-1756: arrays ready to paste into python */
        cout << endl;
        print_times("read_times", read_times);
        print_times("redist_times", redist_times);
        print_times("sort_times", sort_times);
        print_times("cutout_times", cutout_times);
        print_times("write_times", write_times);
    }
}

void processLC(string dir_name, vector<string> out_dirs, vector<string> step_strings, vector<float> halo_pos,
               vector<float> halo_props, float boxLength, int myrank, int numranks, bool verbose, bool timeit,
               bool overwrite, bool positionOnly, bool forceWriteProps, bool propsOnly) {
    const vector<float> box_of_halo(halo_pos.size() / 3, boxLength);
    process_halo_list(dir_name, out_dirs, step_strings, halo_pos, halo_props, box_of_halo, myrank, numranks, verbose,
                      timeit, overwrite, positionOnly, forceWriteProps, propsOnly);
}

namespace cchost {
void processLC_boxes(const string& dir_name, const vector<vector<string>>& out_dirs_of_box,
                     const vector<string>& step_strings, const vector<float>& halo_pos, const vector<float>& halo_props,
                     const vector<float>& boxLengths, int myrank, int numranks, bool verbose, bool timeit, bool overwrite,
                     bool positionOnly, bool forceWriteProps, bool propsOnly) {
    const size_t H = halo_pos.size() / 3, B = boxLengths.size();
    if (out_dirs_of_box.size() != B) die("processLC_boxes: one list of output directories per box length is required");
    const size_t P = H ? halo_props.size() / H : 0;
    vector<string> out_dirs;
    vector<float> pos, props, box;
    for (size_t b = 0; b < B; ++b) {
        if (out_dirs_of_box[b].size() != H) die("processLC_boxes: one output directory per halo and box length is required");
        out_dirs.insert(out_dirs.end(), out_dirs_of_box[b].begin(), out_dirs_of_box[b].end());
        pos.insert(pos.end(), halo_pos.begin(), halo_pos.end());
        props.insert(props.end(), halo_props.begin(), halo_props.begin() + (ptrdiff_t)(P * H));
        box.insert(box.end(), H, boxLengths[b]);
    }
    process_halo_list(dir_name, out_dirs, step_strings, pos, props, box, myrank, numranks, verbose, timeit, overwrite,
                      positionOnly, forceWriteProps, propsOnly);
}
}  // namespace cchost
