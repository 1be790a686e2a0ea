/*
 * oracle/mpi_stub/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A tiny thread-backed MPI subset, written from scratch for this repo, so that the UNMODIFIED
 * reference sources under /root/reference (conflux::lu_params, conflux::LU_rep, conflux_layout;
 * see oracle/build_ref.sh) can be compiled and run here without an MPI installation -- including
 * MULTI-RANK grids: every "rank" is a std::thread of one process (see mpi_threads.cpp).
 *
 * Only the calls that appear on the reference's LU path are provided
 * (grep MPI_ in src/conflux/lu/{conflux_opt.hpp,lu_params.hpp,layout.cpp}).
 */
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __cplusplus
extern "C" {
#endif

struct stub_comm;
struct stub_win;
struct stub_req;
typedef struct stub_comm* MPI_Comm;
typedef struct stub_win* MPI_Win;
typedef struct stub_req* MPI_Request;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Group;
typedef int MPI_Info;
typedef std::ptrdiff_t MPI_Aint;
typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; int count_bytes; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_WIN_NULL ((MPI_Win)0)
#define MPI_INFO_NULL 0
#define MPI_GROUP_NULL 0
extern MPI_Comm stub_comm_world(void);
#define MPI_COMM_WORLD (stub_comm_world())
#define MPI_IN_PLACE ((void*)-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_UNDEFINED (-32766)
#define MPI_MODE_NOPRECEDE 1
#define MPI_SUM 1

/* datatypes: value = size in bytes, tagged in the high bits so distinct types stay distinct */
#define STUB_DT(id, sz) (((id) << 8) | (sz))
#define MPI_CHAR STUB_DT(1, 1)
#define MPI_SHORT STUB_DT(2, 2)
#define MPI_INT STUB_DT(3, 4)
#define MPI_UNSIGNED STUB_DT(4, 4)
#define MPI_UINT32_T STUB_DT(5, 4)
#define MPI_FLOAT STUB_DT(6, 4)
#define MPI_DOUBLE STUB_DT(7, 8)
#define MPI_UNSIGNED_LONG STUB_DT(8, 8)
#define MPI_UNSIGNED_LONG_LONG STUB_DT(9, 8)
#define MPI_UNSIGNED_CHAR STUB_DT(10, 1)
#define MPI_CXX_BOOL STUB_DT(11, 1)
#define MPI_CXX_FLOAT_COMPLEX STUB_DT(12, 8)
#define MPI_CXX_DOUBLE_COMPLEX STUB_DT(13, 16)

int MPI_Init(int*, char***);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_free(MPI_Comm*);
int MPI_Barrier(MPI_Comm);
int MPI_Cart_create(MPI_Comm, int ndims, const int* dims, const int* periods, int reorder, MPI_Comm* out);
int MPI_Cart_sub(MPI_Comm, const int* remain, MPI_Comm* out);
int MPI_Cart_coords(MPI_Comm, int rank, int maxdims, int* coords);
int MPI_Cart_rank(MPI_Comm, const int* coords, int* rank);
int MPI_Cart_get(MPI_Comm, int maxdims, int* dims, int* periods, int* coords);
int MPI_Comm_group(MPI_Comm, MPI_Group*);
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_free(MPI_Group*);
int MPI_Comm_create_group(MPI_Comm, MPI_Group, int, MPI_Comm*);

int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype, MPI_Op, int root, MPI_Comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype, int root, MPI_Comm);
int MPI_Ibcast(void* buf, int count, MPI_Datatype, int root, MPI_Comm, MPI_Request*);
int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype, void* rbuf, int rcount, MPI_Datatype, MPI_Comm);
int MPI_Sendrecv(const void* sbuf, int scount, MPI_Datatype, int dest, int stag, void* rbuf, int rcount,
                 MPI_Datatype, int src, int rtag, MPI_Comm, MPI_Status*);
int MPI_Isend(const void* buf, int count, MPI_Datatype, int dest, int tag, MPI_Comm, MPI_Request*);
int MPI_Irecv(void* buf, int count, MPI_Datatype, int src, int tag, MPI_Comm, MPI_Request*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Waitall(int, MPI_Request*, MPI_Status*);
int MPI_Waitany(int, MPI_Request*, int* idx, MPI_Status*);
int MPI_Request_free(MPI_Request*);
int MPI_Get_count(const MPI_Status*, MPI_Datatype, int* count);
int MPI_Iscatterv(const void* sbuf, const int* scounts, const int* displs, MPI_Datatype, void* rbuf, int rcount,
                  MPI_Datatype, int root, MPI_Comm, MPI_Request*);
int MPI_Igather(const void* sbuf, int scount, MPI_Datatype, void* rbuf, int rcount, MPI_Datatype, int root,
                MPI_Comm, MPI_Request*);
int MPI_Igatherv(const void* sbuf, int scount, MPI_Datatype, void* rbuf, const int* rcounts, const int* displs,
                 MPI_Datatype, int root, MPI_Comm, MPI_Request*);

int MPI_Info_create(MPI_Info*);
int MPI_Info_set(MPI_Info, const char*, const char*);
int MPI_Info_free(MPI_Info*);
int MPI_Win_create(void* base, MPI_Aint size, int disp_unit, MPI_Info, MPI_Comm, MPI_Win*);
int MPI_Win_fence(int, MPI_Win);
int MPI_Win_free(MPI_Win*);
int MPI_Put(const void* origin, int ocount, MPI_Datatype, int target_rank, MPI_Aint target_disp, int tcount,
            MPI_Datatype, MPI_Win);

/* launcher: run fn(rank, arg) on nranks threads that share one MPI_COMM_WORLD */
void stub_mpi_run(int nranks, void (*fn)(int rank, void* arg), void* arg);

#ifdef __cplusplus
}
#endif
