/* host/column_reader.cpp -- see column_reader.h */
#include "column_reader.h"

#include <fcntl.h>
#include <stdio.h>
#include <sys/stat.h>
#include <unistd.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <thread>

#include "util.h"

namespace cchost {

void* PinnedBuffer::ensure(size_t bytes) {
    if (bytes <= bytes_ && p_) return p_;
    release();
    const size_t want = bytes < 64 ? 64 : bytes;
    if (cc_host_alloc(&p_, want) != CC_OK) die(std::string("pinned host allocation failed: ") + cc_last_error());
    bytes_ = want;
    return p_;
}

void PinnedBuffer::release() {
    if (p_) cc_host_free(p_);
    p_ = nullptr;
    bytes_ = 0;
}

/* ---------------------------------------------------------------- GenericIO on-disk format (read side)
 * All integers are u64 and all reals f64 in the FILE's byte order ("HACC01L" little, "HACC01B" big):
 *   global header   Magic[8] HeaderSize NElems Dims[3] NVars VarsSize VarsStart NRanks RanksSize RanksStart
 *                   GlobalHeaderSize PhysOrigin[3] PhysScale[3] BlocksSize BlocksStart      (the last two, like any
 *                   later field, only when GlobalHeaderSize says they are there)
 *   variable header Name[256] Flags Size [ElementSize]           VarsSize bytes each, NVars of them at VarsStart
 *   rank header     Coords[3] NElems Start [GlobalRank]          RanksSize bytes each, NRanks of them at RanksStart
 *   block header    Filters[4][8] Start Size                     BlocksSize bytes each, NRanks x NVars at BlocksStart
 *   CRC64 (8 bytes) closes the header (HeaderSize includes it) and follows every variable's data of every rank.
 * Without block headers the variables of a rank follow each other from the rank's Start, each NElems*Size bytes
 * plus the 8-byte CRC. */
namespace {
struct FileBytes {
    int fd = -1;
    uint64_t size = 0;
    explicit FileBytes(const std::string& f) {
        fd = open(f.c_str(), O_RDONLY);
        struct stat st;
        if (fd >= 0 && fstat(fd, &st) == 0) size = (uint64_t)st.st_size;
    }
    ~FileBytes() { if (fd >= 0) close(fd); }
    bool get(uint64_t off, void* dst, size_t n) const {
        if (fd < 0 || off + n > size) return false;
        size_t done = 0;
        while (done < n) {
            const ssize_t got = pread(fd, (char*)dst + done, n - done, (off_t)(off + done));
            if (got <= 0) return false;
            done += (size_t)got;
        }
        return true;
    }
};
inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

/* GenericIO's checksum (its CRC64.h: "ECMA-182 polynomial with -1 initialization, ... the bit-reversed CRC"), i.e.
 * CRC-64/XZ: reflected polynomial 0xC96C5795D7870F42, initial value and final xor ~0 (check value of "123456789":
 * 0x995DC9BBDF1939FA, tests/test_host_cpu.py).  The writer appends to every checksummed region the 8 bytes that make
 * the CRC of region + those bytes equal ~0 (crc64_invert), which is what a reader verifies.  Slice-by-8 tables. */
struct Crc64 {
    uint64_t t[8][256];
    Crc64() {
        const uint64_t poly = 0xC96C5795D7870F42ull;
        for (int i = 0; i < 256; ++i) {
            uint64_t c = (uint64_t)i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ poly : c >> 1;
            t[0][i] = c;
        }
        for (int i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = t[0][t[s - 1][i] & 0xff] ^ (t[s - 1][i] >> 8);
    }
    /* running register (start with ~0; the CRC is ~register) */
    uint64_t update(uint64_t c, const unsigned char* p, size_t n) const {
        while (n && ((uintptr_t)p & 7)) { c = t[0][(c ^ *p++) & 0xff] ^ (c >> 8); --n; }
        while (n >= 8) {
            uint64_t w;
            memcpy(&w, p, 8); /* little-endian host (checked by the caller) */
            c ^= w;
            c = t[7][c & 0xff] ^ t[6][(c >> 8) & 0xff] ^ t[5][(c >> 16) & 0xff] ^ t[4][(c >> 24) & 0xff] ^
                t[3][(c >> 32) & 0xff] ^ t[2][(c >> 40) & 0xff] ^ t[1][(c >> 48) & 0xff] ^ t[0][c >> 56];
            p += 8;
            n -= 8;
        }
        while (n--) c = t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
        return c;
    }
};
const Crc64& crc64_tables() { static const Crc64 c; return c; }
enum CrcMode { CRC_STRICT, CRC_WARN, CRC_OFF };
CrcMode crc_mode() {
    const char* e = getenv("CC_GIO_CRC");
    if (!e || !*e || !strcmp(e, "strict")) return CRC_STRICT;
    if (!strcmp(e, "warn")) return CRC_WARN;
    return CRC_OFF;
}
/* `reg`: CRC register after region + its 8 trailing bytes; a valid region leaves it at 0 (CRC == ~0) */
void crc_verdict(uint64_t reg, const std::string& what, std::string* err) {
    if (reg == 0 || crc_mode() == CRC_OFF) return;
    const std::string msg = "CRC64 mismatch in " + what + " (the file is corrupt or truncated; CC_GIO_CRC=warn reads it anyway, "
                            "CC_GIO_CRC=off skips the check)";
    if (crc_mode() == CRC_WARN) { fprintf(stdout, "WARNING: GenericIO reader: %s\n", msg.c_str()); return; }
    if (err) *err = msg;
    else cchost::die("GenericIO reader: " + msg);
}
inline void swap_elems(void* p, size_t n, size_t size) {
    if (size == 2) { uint16_t* q = (uint16_t*)p; for (size_t i = 0; i < n; ++i) q[i] = __builtin_bswap16(q[i]); }
    else if (size == 4) { uint32_t* q = (uint32_t*)p; for (size_t i = 0; i < n; ++i) q[i] = __builtin_bswap32(q[i]); }
    else if (size == 8) { uint64_t* q = (uint64_t*)p; for (size_t i = 0; i < n; ++i) q[i] = __builtin_bswap64(q[i]); }
}
bool host_is_little() { const uint16_t one = 1; return *(const unsigned char*)&one == 1; }
bool is_gio_magic(const char* m, bool* big) {
    if (memcmp(m, "HACC01L", 8) == 0) { *big = false; return true; }
    if (memcmp(m, "HACC01B", 8) == 0) { *big = true; return true; }
    return false;
}
}  // namespace

void ColumnReader::parseGio(const std::string& file, bool is_main, std::vector<int32_t>* partitions) {
    FileBytes fb(file);
    char magic[8];
    bool big = false;
    if (!fb.get(0, magic, 8) || !is_gio_magic(magic, &big)) die("GenericIO reader: " + file + " is not a GenericIO file");
    const bool swap = big == host_is_little();
    auto u64_at = [&](uint64_t off) {
        uint64_t v = 0;
        if (!fb.get(off, &v, 8)) die("GenericIO reader: truncated header in " + file);
        return swap ? bswap64(v) : v;
    };
    /* field k (0-based, 8 bytes each) after the magic */
    const uint64_t header_size = u64_at(8 + 8 * 0), nvars = u64_at(8 + 8 * 5), vars_size = u64_at(8 + 8 * 6),
                   vars_start = u64_at(8 + 8 * 7), nranks = u64_at(8 + 8 * 8), ranks_size = u64_at(8 + 8 * 9),
                   ranks_start = u64_at(8 + 8 * 10), global_size = u64_at(8 + 8 * 11);
    /* PhysOrigin[3], PhysScale[3] are fields 12..17; BlocksSize, BlocksStart 18, 19 */
    uint64_t blocks_size = 0, blocks_start = 0;
    if (global_size >= 8 + 8 * 20) { blocks_size = u64_at(8 + 8 * 18); blocks_start = u64_at(8 + 8 * 19); }
    if (header_size > fb.size || header_size < 8 + 8 * 12 + 8 || vars_size < 256 + 16 || ranks_size < 40 || nvars > 4096 ||
        nranks > (1u << 24) || vars_start + nvars * vars_size > header_size || ranks_start + nranks * ranks_size > header_size)
        die("GenericIO reader: implausible header in " + file);
    if (crc_mode() != CRC_OFF) { /* HeaderSize counts the header's own 8 CRC bytes */
        std::vector<unsigned char> hdr((size_t)header_size);
        if (!fb.get(0, hdr.data(), hdr.size())) die("GenericIO reader: truncated header in " + file);
        crc_verdict(crc64_tables().update(~0ull, hdr.data(), hdr.size()), "the header of " + file, nullptr);
    }
    std::vector<GioVar> vars((size_t)nvars);
    for (uint64_t v = 0; v < nvars; ++v) {
        char name[257];
        if (!fb.get(vars_start + v * vars_size, name, 256)) die("GenericIO reader: truncated variable table in " + file);
        name[256] = 0;
        vars[(size_t)v].name = name;
        vars[(size_t)v].size = u64_at(vars_start + v * vars_size + 256 + 8);
    }
    const bool has_blocks = blocks_size >= 48 && blocks_start != 0;
    for (uint64_t r = 0; r < nranks; ++r) {
        const uint64_t o = ranks_start + r * ranks_size;
        GioBlock b;
        b.file = file;
        b.swap = swap;
        b.nelems = u64_at(o + 24);
        b.start = u64_at(o + 32);
        b.global_rank = ranks_size >= 48 ? u64_at(o + 40) : r;
        b.vars = vars;
        b.first_elem = 0;
        if (has_blocks) {
            for (uint64_t v = 0; v < nvars; ++v) {
                const uint64_t bo = blocks_start + (r * nvars + v) * blocks_size;
                char filt[9];
                if (!fb.get(bo, filt, 8)) die("GenericIO reader: truncated block table in " + file);
                filt[8] = 0;
                if (filt[0] != 0) /* fail before anything is read or allocated: there is no decoder here */
                    die("GenericIO reader: variable '" + vars[(size_t)v].name + "' of " + file + " is stored through the '" + filt +
                        "' filter (compressed GenericIO output).  This build reads uncompressed blocks only: re-write the "
                        "lightcone with compression off (GENERICIO_COMPRESS=0) or decompress it with GenericIO's own tools.");
                b.var_filtered.push_back(false);
                b.var_start.push_back(u64_at(bo + 32));
                b.var_bytes.push_back(u64_at(bo + 40));
            }
        }
        if (is_main && partitions) {
            /* a partitioned main file holds only the rank map: its own blocks carry "$rank"/"$partition" */
            bool map_only = false;
            for (const GioVar& gv : vars) map_only = map_only || gv.name == "$partition";
            if (map_only) {
                /* read this block's $partition values */
                uint64_t off = b.start;
                for (size_t v = 0; v < vars.size(); ++v) {
                    const uint64_t at = has_blocks ? b.var_start[v] : off;
                    if (vars[v].name == "$partition") {
                        if (vars[v].size != 4) die("GenericIO reader: $partition is not 32-bit in " + file);
                        std::vector<int32_t> part((size_t)b.nelems);
                        if (b.nelems && !fb.get(at, part.data(), (size_t)b.nelems * 4)) die("GenericIO reader: truncated rank map in " + file);
                        if (swap) swap_elems(part.data(), part.size(), 4);
                        partitions->insert(partitions->end(), part.begin(), part.end());
                    }
                    off += b.nelems * vars[v].size + 8;
                }
                continue;
            }
        }
        blocks_.push_back(b);
    }
}

void ColumnReader::openAndReadHeader() {
    blocks_.clear();
    gio_ = false;
    {
        FileBytes fb(path_);
        char magic[8];
        bool big;
        gio_ = fb.get(0, magic, 8) && is_gio_magic(magic, &big);
    }
    if (!gio_) {
        struct stat st;
        const std::string xfile = path_ + "#x";
        if (stat(xfile.c_str(), &st) != 0)
            die("column reader: " + path_ + " is not a GenericIO file and there is no column store " + xfile +
                " (one raw little-endian file per variable next to the header)");
        nelems_ = (size_t)st.st_size / 4;
        return;
    }
    std::vector<int32_t> partitions;
    parseGio(path_, true, &partitions);
    if (!partitions.empty()) { /* partitioned: the data lives in <header>#<p>, each a complete GenericIO file */
        std::vector<int32_t> uniq(partitions);
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        blocks_.clear();
        for (int32_t p : uniq) parseGio(path_ + "#" + std::to_string(p), false, nullptr);
    }
    /* element order of a single-rank read: writing ranks in global rank order */
    std::stable_sort(blocks_.begin(), blocks_.end(), [](const GioBlock& a, const GioBlock& b) { return a.global_rank < b.global_rank; });
    uint64_t total = 0;
    for (GioBlock& b : blocks_) { b.first_elem = total; total += b.nelems; }
    nelems_ = (size_t)total;
}

void ColumnReader::addVariable(const std::string& name, PinnedBuffer* dst, size_t elem_size) {
    vars_.push_back(Var{name, dst, elem_size});
}

static void read_whole(const std::string& file, void* dst, size_t bytes, std::string* err) {
    const int fd = open(file.c_str(), O_RDONLY);
    if (fd < 0) { *err = "cannot open " + file; return; }
    size_t done = 0;
    while (done < bytes) {
        const ssize_t got = pread(fd, (char*)dst + done, bytes - done, (off_t)done);
        if (got <= 0) { *err = "short read on " + file; break; }
        done += (size_t)got;
    }
    close(fd);
}

void ColumnReader::readGioVar(const Var& v, std::string* err) const {
    char* dst = (char*)v.dst->data();
    for (const GioBlock& b : blocks_) {
        size_t vi = b.vars.size();
        uint64_t off = b.start;
        for (size_t i = 0; i < b.vars.size(); ++i) {
            if (b.vars[i].name == v.name) { vi = i; break; }
            off += b.nelems * b.vars[i].size + 8; /* data + CRC64 */
        }
        if (vi == b.vars.size()) { *err = "variable " + v.name + " not in " + b.file; return; }
        if (b.vars[vi].size != v.elem_size) {
            *err = "variable " + v.name + " has " + std::to_string(b.vars[vi].size) + "-byte elements in " + b.file + ", expected " +
                   std::to_string(v.elem_size);
            return;
        }
        if (!b.var_start.empty()) {
            if (b.var_filtered[vi]) { *err = "variable " + v.name + " in " + b.file + " is compressed (filter blocks are not supported)"; return; }
            off = b.var_start[vi];
        }
        if (b.nelems == 0) continue;
        FileBytes fb(b.file);
        char* at = dst + b.first_elem * v.elem_size;
        const size_t nbytes = (size_t)b.nelems * v.elem_size;
        if (!fb.get(off, at, nbytes)) { *err = "short read of " + v.name + " in " + b.file; return; }
        if (crc_mode() != CRC_OFF) { /* the block's data is followed by the 8 bytes that close its CRC */
            unsigned char tail[8];
            if (!fb.get(off + nbytes, tail, 8)) { *err = "no CRC behind " + v.name + " in " + b.file; return; }
            uint64_t reg = crc64_tables().update(~0ull, (const unsigned char*)at, nbytes);
            reg = crc64_tables().update(reg, tail, 8);
            crc_verdict(reg, "variable " + v.name + ", rank block " + std::to_string(b.global_rank) + " of " + b.file, err);
            if (!err->empty()) return;
        }
        if (b.swap) swap_elems(at, (size_t)b.nelems, v.elem_size);
    }
}

void ColumnReader::allocate() {
    for (const Var& v : vars_) v.dst->ensure(nelems_ * v.elem_size);
}

void ColumnReader::readData() {
    std::vector<std::string> errs(vars_.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < vars_.size(); ++i) {
        void* dst = vars_[i].dst->data();
        if (gio_) th.emplace_back(&ColumnReader::readGioVar, this, std::cref(vars_[i]), &errs[i]);
        else th.emplace_back(read_whole, path_ + "#" + vars_[i].name, dst, nelems_ * vars_[i].elem_size, &errs[i]);
    }
    for (auto& t : th) t.join();
    for (const auto& e : errs)
        if (!e.empty()) die("column reader: " + e);
}

void StepColumns::prepare(const std::string& header_path, bool positionOnly) {
    reader_.clear();
    reader_.emplace_back(header_path);
    ColumnReader& rd = reader_[0];
    rd.openAndReadHeader();
    n = rd.readNumElems();
    position_only = positionOnly;
    rd.addVariable("x", &x, 4);
    rd.addVariable("y", &y, 4);
    rd.addVariable("z", &z, 4);
    rd.addVariable("a", &a, 4);
    rd.addVariable("id", &id, 8);
    if (!positionOnly) {
        rd.addVariable("vx", &vx, 4);
        rd.addVariable("vy", &vy, 4);
        rd.addVariable("vz", &vz, 4);
        rd.addVariable("rotation", &rotation, 4);
        rd.addVariable("replication", &replication, 4);
    }
    rd.allocate();
}

void StepColumns::load() {
    if (reader_.empty()) die("column reader: load() without prepare()");
    reader_[0].readData();
    reader_.clear();
}

cc_cols_in StepColumns::view(size_t first) const {
    cc_cols_in c;
    c.x = (const float*)x.data() + first;
    c.y = (const float*)y.data() + first;
    c.z = (const float*)z.data() + first;
    c.a = (const float*)a.data() + first;
    c.id = (const int64_t*)id.data() + first;
    if (position_only) {
        c.vx = c.vy = c.vz = nullptr;
        c.rotation = c.replication = nullptr;
    } else {
        c.vx = (const float*)vx.data() + first;
        c.vy = (const float*)vy.data() + first;
        c.vz = (const float*)vz.data() + first;
        c.rotation = (const int32_t*)rotation.data() + first;
        c.replication = (const int32_t*)replication.data() + first;
    }
    return c;
}

}  // namespace cchost
