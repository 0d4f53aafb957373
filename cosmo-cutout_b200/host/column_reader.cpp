/* host/column_reader.cpp -- see column_reader.h */
#include "column_reader.h"

#include <fcntl.h>
#include <stdio.h>
#include <sys/stat.h>
#include <unistd.h>

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

void ColumnReader::openAndReadHeader() {
    struct stat st;
    const std::string xfile = path_ + "#x";
    if (stat(xfile.c_str(), &st) != 0)
        die("column reader: no column store for " + xfile +
            " (expected one raw little-endian file per variable next to the header)");
    nelems_ = (size_t)st.st_size / 4;
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

void ColumnReader::readData() {
    std::vector<std::string> errs(vars_.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < vars_.size(); ++i) {
        void* dst = vars_[i].dst->ensure(nelems_ * vars_[i].elem_size);
        th.emplace_back(read_whole, path_ + "#" + vars_[i].name, dst, nelems_ * vars_[i].elem_size, &errs[i]);
    }
    for (auto& t : th) t.join();
    for (const auto& e : errs)
        if (!e.empty()) die("column reader: " + e);
}

void StepColumns::read(const std::string& header_path, bool positionOnly) {
    ColumnReader rd(header_path);
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
    rd.readData();
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
