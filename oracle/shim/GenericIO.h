/*
 * oracle/shim/GenericIO.h -- column reader stand-in for ANL/HACC GenericIO, TEST INFRASTRUCTURE ONLY.
 *
 * GenericIO is an un-vendored, un-pinned dependency of the reference (Makefile.cooley:8-11) and is
 * not in this image.  It does I/O only; no hot-path arithmetic lives in it.  This header provides the
 * six-call subset the reference uses (processLC.cpp:95-138, :854-888) over either
 *   (a) columns registered in memory with gio_shim_register(), or
 *   (b) one raw little-endian file per variable named "<header>#<var>" next to an (empty) header
 *       file -- names containing '#' are ignored by the reference's getLCFile (util.cpp:289-291).
 * The same flat layout is what the product's own column reader accepts ("input format").
 */
#ifndef ORACLE_SHIM_GENERICIO_H
#define ORACLE_SHIM_GENERICIO_H

#include <mpi.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

extern "C" {
/* register an in-memory column for "<header path>#<var>"; data is NOT copied. nbytes = total bytes. */
void gio_shim_register(const char* key, const void* data, size_t nbytes);
void gio_shim_clear(void);
/* returns 1 and fills ptr/nbytes when key is registered */
int gio_shim_lookup(const char* key, const void** ptr, size_t* nbytes);
}

namespace gio {

class GenericIO {
public:
    enum FileIO { FileIOMPI, FileIOPOSIX, FileIOMPICollective };
    enum MismatchBehavior { MismatchAllowed, MismatchDisallowed, MismatchRedistribute };

    GenericIO(MPI_Comm, const std::string& fname, unsigned = FileIOPOSIX) : path_(fname), nelems_(0) {}

    void openAndReadHeader(MismatchBehavior = MismatchDisallowed) {
        /* element count comes from the x column: 4 bytes per element */
        size_t nb = 0;
        if (!column_bytes("x", &nb)) {
            fprintf(stderr, "GenericIO shim: no column store for %s#x\n", path_.c_str());
            MPI_Abort(MPI_COMM_WORLD, 1);
        }
        nelems_ = nb / 4;
    }
    size_t readNumElems(int = -1) const { return nelems_; }
    /* the real library asks for room for a trailing CRC64 */
    size_t requestedExtraSpace() const { return 8; }

    template <typename T>
    void addVariable(const std::string& name, std::vector<T>& v, bool = false) {
        Var q;
        q.name = name;
        q.dst = (void*)&v[0];
        q.elsize = sizeof(T);
        vars_.push_back(q);
    }

    void readData() {
        for (size_t i = 0; i < vars_.size(); ++i) {
            const Var& q = vars_[i];
            const size_t want = nelems_ * q.elsize;
            const std::string key = path_ + "#" + q.name;
            const void* src = 0;
            size_t nb = 0;
            if (gio_shim_lookup(key.c_str(), &src, &nb)) {
                if (nb < want) { fprintf(stderr, "GenericIO shim: %s too short\n", key.c_str()); MPI_Abort(MPI_COMM_WORLD, 1); }
                memcpy(q.dst, src, want);
                continue;
            }
            FILE* f = fopen(key.c_str(), "rb");
            if (!f) { fprintf(stderr, "GenericIO shim: cannot open %s\n", key.c_str()); MPI_Abort(MPI_COMM_WORLD, 1); }
            size_t got = fread(q.dst, 1, want, f);
            fclose(f);
            if (got != want) { fprintf(stderr, "GenericIO shim: short read on %s\n", key.c_str()); MPI_Abort(MPI_COMM_WORLD, 1); }
        }
        shim_trace_event("GIO_readData_done");
    }

private:
    struct Var { std::string name; void* dst; size_t elsize; };

    bool column_bytes(const char* var, size_t* nb) const {
        const std::string key = path_ + "#" + var;
        const void* p = 0;
        if (gio_shim_lookup(key.c_str(), &p, nb)) return true;
        FILE* f = fopen(key.c_str(), "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END);
        *nb = (size_t)ftell(f);
        fclose(f);
        return true;
    }

    std::string path_;
    size_t nelems_;
    std::vector<Var> vars_;
};

}  // namespace gio

#endif
