/*
 * host/column_reader.h -- the input side of a lightcone step: named columns -> pinned host buffers.
 *
 * Stands where the reference uses GenericIO (processLC.cpp:94-139, :854-889): open a step's header file, ask for
 * the element count, register the variables wanted, read.  GenericIO itself is an un-vendored, un-pinned external
 * library that is not available here, so the only back-end is the flat column layout
 *     <dir>/<header>            (may be empty)
 *     <dir>/<header>#<var>      raw little-endian array of the variable, one file per variable
 * -- names containing '#' are ignored by header discovery, exactly as GenericIO's own rank files are
 * (util.cpp:289-291).  Buffers are cudaHostAlloc'ed (pinned, mapped) through the C ABI so the GPU can stream
 * x,y,z from them with cudaMemcpyAsync and gather the survivors' other columns in place.
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
    explicit ColumnReader(const std::string& header_path) : path_(header_path), nelems_(0) {}
    /* element count from the size of the x column */
    void openAndReadHeader();
    size_t readNumElems() const { return nelems_; }
    void addVariable(const std::string& name, PinnedBuffer* dst, size_t elem_size);
    /* reads every registered variable (one thread per variable) */
    void readData();

private:
    struct Var { std::string name; PinnedBuffer* dst; size_t elem_size; };
    std::string path_;
    size_t nelems_;
    std::vector<Var> vars_;
};

/* the ten input columns of a step in pinned memory (util.h:33-47 Buffers_read) */
struct StepColumns {
    PinnedBuffer x, y, z, vx, vy, vz, a, id, rotation, replication;
    size_t n = 0;
    bool position_only = false;
    /* processLC.cpp:102-139 (all ten) / :861-889 (five when positionOnly) */
    void read(const std::string& header_path, bool positionOnly);
    cc_cols_in view(size_t first) const;
};

}  // namespace cchost
#endif
