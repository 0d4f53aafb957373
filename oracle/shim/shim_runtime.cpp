/*
 * oracle/shim/shim_runtime.cpp -- runtime for the single-rank MPI / GenericIO stand-ins.
 * TEST INFRASTRUCTURE ONLY (see mpi.h, GenericIO.h in this directory).
 */
#include "mpi.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

struct shim_mpi_file {
    int fd;
    long long pos;
};

namespace {
int g_abort_throws = 0;
struct TraceEv { std::string name; double t; long long value; };
std::vector<TraceEv> g_trace;
std::map<std::string, std::pair<const void*, size_t> > g_columns;

double now_s() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
}  // namespace

extern "C" {

void shim_set_abort_throws(int on) { g_abort_throws = on; }
void shim_trace_reset(void) { g_trace.clear(); }
void shim_trace_event(const char* name) { TraceEv e; e.name = name; e.t = now_s(); e.value = 0; g_trace.push_back(e); }
void shim_trace_event_value(const char* name, long long value) { TraceEv e; e.name = name; e.t = now_s(); e.value = value; g_trace.push_back(e); }
int shim_trace_count(void) { return (int)g_trace.size(); }
const char* shim_trace_name(int i) { return g_trace[i].name.c_str(); }
double shim_trace_time(int i) { return g_trace[i].t; }
long long shim_trace_value(int i) { return g_trace[i].value; }

void gio_shim_register(const char* key, const void* data, size_t nbytes) {
    g_columns[std::string(key)] = std::make_pair(data, nbytes);
}
void gio_shim_clear(void) { g_columns.clear(); }
int gio_shim_lookup(const char* key, const void** ptr, size_t* nbytes) {
    std::map<std::string, std::pair<const void*, size_t> >::const_iterator it = g_columns.find(std::string(key));
    if (it == g_columns.end()) return 0;
    *ptr = it->second.first;
    *nbytes = it->second.second;
    return 1;
}

int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int errorcode) {
    fflush(stdout);
    if (g_abort_throws) {
        shim_abort e;
        e.code = errorcode;
        throw e;
    }
    exit(errorcode ? errorcode : 1);
}
int MPI_Barrier(MPI_Comm) {
    shim_trace_event("MPI_Barrier");
    return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
double MPI_Wtime(void) { return now_s(); }
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }

int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int, MPI_Datatype, MPI_Comm) {
    /* the first element of the send buffer is recorded: use case 2 gathers cutout_size then
     * rough_cutout_size (processLC.cpp:1569-1572), which the reference otherwise only prints */
    long long v = 0;
    if (sdt == 4) v = *(const int*)sbuf; else if (sdt == 8) v = *(const long long*)sbuf;
    shim_trace_event_value("MPI_Allgather", v);
    memmove(rbuf, sbuf, (size_t)scount * (size_t)sdt);
    return MPI_SUCCESS;
}
int MPI_Alltoall(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int, MPI_Datatype, MPI_Comm) {
    memmove(rbuf, sbuf, (size_t)scount * (size_t)sdt);
    return MPI_SUCCESS;
}
int MPI_Alltoallv(const void* sbuf, const int* scounts, const int* sdispls, MPI_Datatype sdt, void* rbuf,
                  const int*, const int* rdispls, MPI_Datatype rdt, MPI_Comm) {
    /* one rank: block 0 goes to block 0 */
    memmove((char*)rbuf + (size_t)rdispls[0] * (size_t)rdt, (const char*)sbuf + (size_t)sdispls[0] * (size_t)sdt,
            (size_t)scounts[0] * (size_t)sdt);
    return MPI_SUCCESS;
}
int MPI_Type_struct(int, int*, MPI_Aint*, MPI_Datatype*, MPI_Datatype* newtype) { *newtype = 0; return MPI_SUCCESS; }
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }

int MPI_File_open(MPI_Comm, char* name, int, MPI_Info, MPI_File* fh) {
    /* MPI_MODE_CREATE|MPI_MODE_WRONLY: create, do not truncate */
    shim_mpi_file* f = new shim_mpi_file;
    f->fd = open(name, O_CREAT | O_WRONLY, 0644);
    f->pos = 0;
    if (f->fd < 0) {
        fprintf(stderr, "MPI shim: cannot open %s\n", name);
        delete f;
        return MPI_Abort(MPI_COMM_WORLD, 1);
    }
    *fh = f;
    shim_trace_event("MPI_File_open");
    return MPI_SUCCESS;
}
int MPI_File_close(MPI_File* fh) {
    close((*fh)->fd);
    delete *fh;
    *fh = 0;
    return MPI_SUCCESS;
}
int MPI_File_seek(MPI_File fh, MPI_Offset off, int) { fh->pos = off; return MPI_SUCCESS; }
int MPI_File_iwrite(MPI_File fh, const void* buf, int count, MPI_Datatype dt, MPI_Request* req) {
    size_t n = (size_t)count * (size_t)dt;
    const char* p = (const char*)buf;
    long long pos = fh->pos;
    while (n > 0) {
        ssize_t w = pwrite(fh->fd, p, n, pos);
        if (w <= 0) { fprintf(stderr, "MPI shim: write failed\n"); return MPI_Abort(MPI_COMM_WORLD, 1); }
        n -= (size_t)w; p += w; pos += w;
    }
    fh->pos = pos;
    if (req) *req = 0;
    return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }

}  // extern "C"
