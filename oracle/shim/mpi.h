/*
 * oracle/shim/mpi.h -- single-rank MPI stand-in, TEST INFRASTRUCTURE ONLY.
 *
 * The reference (jhollowed/cosmo-cutout) needs <mpi.h>; no MPI exists in this image.
 * This header gives the 18 MPI entry points the reference calls the
 * semantics they have on a communicator of size 1, so that /root/reference/src/*.cpp can be
 * compiled where they lie (oracle/Makefile) into oracle/_ref/.  Nothing under
 * cosmo-cutout_b200/ may include this file.
 *
 * Besides the MPI semantics, every collective / file call is time-stamped into a global trace
 * (shim_trace_*) so a harness can time the reference's hot loop without modifying it: in use
 * case 1 the loop (processLC.cpp:276-310) is exactly the interval between the last
 * MPI_File_open (processLC.cpp:263) and the next MPI_Barrier (processLC.cpp:311).
 */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H

#include <stddef.h>
#include <stdint.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* value = element size in bytes */
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef int MPI_Info;
typedef int MPI_Request;
typedef struct { int unused; } MPI_Status;
typedef struct shim_mpi_file* MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_SEEK_SET 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 4
#define MPI_SUCCESS 0

#define MPI_CHAR 1
#define MPI_FLOAT 4
#define MPI_INT 4
#define MPI_INT32_T 4
#define MPI_INT64_T 8
#define MPI_DOUBLE 8

#ifdef __cplusplus
extern "C" {
#endif

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Barrier(MPI_Comm comm);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
double MPI_Wtime(void);
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount,
                  MPI_Datatype rdt, MPI_Comm comm);
int MPI_Alltoall(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount,
                 MPI_Datatype rdt, MPI_Comm comm);
int MPI_Alltoallv(const void* sbuf, const int* scounts, const int* sdispls, MPI_Datatype sdt,
                  void* rbuf, const int* rcounts, const int* rdispls, MPI_Datatype rdt,
                  MPI_Comm comm);
int MPI_Type_struct(int count, int* blocklens, MPI_Aint* disps, MPI_Datatype* types,
                    MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* dt);
int MPI_File_open(MPI_Comm comm, char* name, int amode, MPI_Info info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence);
int MPI_File_iwrite(MPI_File fh, const void* buf, int count, MPI_Datatype dt, MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* st);

/* ---- harness hooks (not MPI) ---- */
/* When non-zero, MPI_Abort throws a C++ exception (shim_abort) instead of exiting the process. */
void shim_set_abort_throws(int on);
void shim_trace_reset(void);
void shim_trace_event(const char* name);
void shim_trace_event_value(const char* name, long long value);
int shim_trace_count(void);
const char* shim_trace_name(int i);
double shim_trace_time(int i);
long long shim_trace_value(int i);

#ifdef __cplusplus
}
struct shim_abort { int code; };
#endif

#endif
