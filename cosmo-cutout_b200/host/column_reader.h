/*
 * host/column_reader.h -- the input side of a lightcone step: named columns -> pinned host buffers.
 *
 * Stands where the reference uses GenericIO (processLC.cpp:94-139, :854-889): open a step's header file, ask for
 * the element count, register the variables wanted, read.  Two back-ends, chosen by the header file's first bytes:
 *
 *   flat columns     <dir>/<header>            (empty, or anything that is not a GenericIO file)
 *                    <dir>/<header>#<var>      raw little-endian array of the variable, one file per variable
 *   GenericIO file   <dir>/<header> starts with the magic "HACC01L" / "HACC01B": the self-describing HACC format
 *                    (global header, variable headers, one header per writing rank, optional block headers, CRC64
 *                    after the header and after every data block), optionally partitioned into <header>#<p>
 *                    files through the "$rank"/"$partition" map of the main file.  Ranks are read in global rank
 *                    order, which is the element order a single-rank GenericIO::readData (MismatchRedistribute)
 *                    produces.  Both byte orders.  The CRC64 behind the header and behind every variable block is
 *                    verified (CRC-64/XZ, GenericIO's CRC64.h; a mismatch is fatal, CC_GIO_CRC=warn|off relaxes it).
 *                    Uncompressed blocks only: a file written through a compression filter (blosc) is refused with
 *                    an explanation when its header is parsed.
 *
 * GenericIO itself is an un-vendored, un-pinned external library that is not available here: the
 * second back-end is written from the published on-disk format and is exercised against files produced by an
 * independent writer in tests/harness.py (write_gio_step), NOT against the library -- treat it as unverified on
 * real HACC output until it has read one.  Names containing '#' are ignored by header discovery, exactly as the
 * reference ignores GenericIO's partition files (util.cpp:289-291).  Buffers are cudaHostAlloc'ed (pinned, mapped)
 * through the C ABI so the GPU can stream x,y,z from them with cudaMemcpyAsync.
 */
#ifndef CC_HOST_COLUMN_READER_H
#define CC_HOST_COLUMN_READER_H

#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "cosmo_cutout_b200.h"

namespace cchost {

class PinnedBuffer {
public:
    PinnedBuffer() : p_(nullptr), bytes_(0) {}
    ~PinnedBuffer() { release(); }
    PinnedBuffer(const PinnedBuffer&) = delete;
    PinnedBuffer& operator=(const PinnedBuffer&) = delete;
    void* ensure(size_t bytes); /* grows, never shrinks; dies on failure */
    void release();
    void* data() const { return p_; }

private:
    void* p_;
    size_t bytes_;
};

class ColumnReader {
public:
    explicit ColumnReader(const std::string& header_path) : path_(header_path), nelems_(0), gio_(false) {}
    /* flat columns: element count from the size of the x column; GenericIO: sum over the rank headers */
    void openAndReadHeader();
    bool isGenericIO() const { return gio_; }
    size_t readNumElems() const { return nelems_; }
    void addVariable(const std::string& name, PinnedBuffer* dst, size_t elem_size);
    /* sizes the destination buffers (pinned allocation) */
    void allocate();
    /* reads every registered variable (one thread per variable); buffers must have been allocated */
    void readData();

private:
    struct Var { std::string name; PinnedBuffer* dst; size_t elem_size; };
    /* one writing rank's block of one GenericIO (partition) file */
    struct GioVar { std::string name; uint64_t size; };
    struct GioBlock {
        std::string file;
        bool swap;                      /* file byte order differs from the host's */
        uint64_t global_rank, nelems, start;
        std::vector<GioVar> vars;       /* in file order */
        std::vector<uint64_t> var_start; /* explicit per-variable offsets (block headers) or empty */
        std::vector<bool> var_filtered;  /* block carries a compression filter (refused when the header is parsed) */
        std::vector<uint64_t> var_bytes; /* block header: stored size of the variable's data */
        uint64_t first_elem;            /* position of the block's first element in the assembled column */
    };
    void parseGio(const std::string& file, bool is_main, std::vector<int32_t>* partitions);
    void readGioVar(const Var& v, std::string* err) const;

    std::string path_;
    size_t nelems_;
    bool gio_;
    std::vector<Var> vars_;
    std::vector<GioBlock> blocks_;
};

/* the ten input columns of a step in pinned memory (util.h:33-47 Buffers_read) */
struct StepColumns {
    PinnedBuffer x, y, z, vx, vy, vz, a, id, rotation, replication;
    size_t n = 0;
    bool position_only = false;
    /* processLC.cpp:102-139 (all ten) / :861-889 (five when positionOnly) */
    void read(const std::string& header_path, bool positionOnly) { prepare(header_path, positionOnly); load(); }
    /* read = prepare (open, parse the header, size the pinned buffers: may call cudaHostAlloc / cudaFreeHost, so only
     * on a thread and at a time where no GPU work waits on another GPU) + load (file I/O only: any thread) */
    void prepare(const std::string& header_path, bool positionOnly);
    void load();
    cc_cols_in view(size_t first) const;

private:
    std::vector<ColumnReader> reader_; /* 0 or 1: the prepared read */
};

}  // namespace cchost
#endif
