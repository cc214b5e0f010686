/*
 * mpi.h — minimpi: a small single-node MPI subset over POSIX shared memory.
 *
 * TEST / BASELINE INFRASTRUCTURE.  There is no MPI in this image, and the reference (hypre
 * 3.1.0) only runs on more than one rank through MPI: its multi-rank BoomerAMG setup (whose
 * hierarchy the B200 solve consumes) and its multi-rank CPU solve (the oracle / CPU baseline)
 * need one.  minimpi implements exactly the calls hypre's wrappers use
 * (src/utilities/mpistubs.c:940-1793 of the reference): point-to-point with tag matching,
 * MPI_ANY_SOURCE / MPI_ANY_TAG, probe, persistent requests, the usual collectives,
 * communicator / group management and derived datatypes.  Ranks are processes of one node,
 * started by oracle/minimpi/mpirun or by torchrun (RANK / WORLD_SIZE / MASTER_PORT).
 * Not a product component: the GPU solve path communicates through NCCL only.
 */
#ifndef MINIMPI_H
#define MINIMPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPI_VERSION 3
#define MPI_SUBVERSION 1

typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Request;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef int MPI_Fint;

typedef struct
{
   int MPI_SOURCE;
   int MPI_TAG;
   int MPI_ERROR;
   long long _nbytes;
} MPI_Status;

typedef void (MPI_User_function)(void *invec, void *inoutvec, int *len, MPI_Datatype *datatype);

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15

#define MPI_COMM_NULL  0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF  2
#define MPI_GROUP_NULL 0
#define MPI_REQUEST_NULL 0
#define MPI_INFO_NULL 0
#define MPI_COMM_TYPE_SHARED 1

#define MPI_BOTTOM ((void *) 0)
#define MPI_STATUS_IGNORE   ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG    (-1)
#define MPI_UNDEFINED  (-32766)
#define MPI_PROC_NULL  (-2)

/* builtin datatypes (ids < 32) */
#define MPI_DATATYPE_NULL 0
#define MPI_CHAR          1
#define MPI_BYTE          2
#define MPI_SHORT         3
#define MPI_INT           4
#define MPI_LONG          5
#define MPI_LONG_LONG_INT 6
#define MPI_LONG_LONG     6
#define MPI_UNSIGNED      7
#define MPI_FLOAT         8
#define MPI_DOUBLE        9
#define MPI_LONG_DOUBLE   10
#define MPI_C_FLOAT_COMPLEX       11
#define MPI_C_DOUBLE_COMPLEX      12
#define MPI_C_LONG_DOUBLE_COMPLEX 13
#define MPI_UNSIGNED_LONG 14
#define MPI_REAL          8

/* reduction ops (ids < 16; user ops above) */
#define MPI_OP_NULL 0
#define MPI_SUM  1
#define MPI_MIN  2
#define MPI_MAX  3
#define MPI_LOR  4
#define MPI_LAND 5
#define MPI_BOR  6
#define MPI_BAND 7
#define MPI_PROD 8

int MPI_Init(int *argc, char ***argv);
int MPI_Initialized(int *flag);
int MPI_Finalize(void);
int MPI_Finalized(int *flag);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);
double MPI_Wtick(void);

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);
int MPI_Comm_split_type(MPI_Comm comm, int split_type, int key, MPI_Info info, MPI_Comm *newcomm);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup);
int MPI_Group_free(MPI_Group *group);
MPI_Comm MPI_Comm_f2c(MPI_Fint comm);
MPI_Fint MPI_Comm_c2f(MPI_Comm comm);
int MPI_Info_create(MPI_Info *info);
int MPI_Info_free(MPI_Info *info);

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irsend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Send_init(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Recv_init(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Start(MPI_Request *req);
int MPI_Startall(int count, MPI_Request *reqs);
int MPI_Probe(int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int *flag, MPI_Status *status);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status);
int MPI_Testall(int count, MPI_Request *reqs, int *flag, MPI_Status *statuses);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *statuses);
int MPI_Waitany(int count, MPI_Request *reqs, int *index, MPI_Status *status);
int MPI_Request_free(MPI_Request *req);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype dt, int *count);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Scan(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                const int *recvcounts, const int *displs, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Scatter(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype sendtype,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                   const int *recvcounts, const int *displs, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoall(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                 MPI_Datatype recvtype, MPI_Comm comm);

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_create_hvector(int count, int blocklength, MPI_Aint stride, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_hvector(int count, int blocklength, MPI_Aint stride, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_create_struct(int count, const int *blocklengths, const MPI_Aint *displs,
                           const MPI_Datatype *types, MPI_Datatype *newtype);
int MPI_Type_struct(int count, int *blocklengths, MPI_Aint *displs, MPI_Datatype *types, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *dt);
int MPI_Type_free(MPI_Datatype *dt);
int MPI_Type_size(MPI_Datatype dt, int *size);
int MPI_Get_address(const void *location, MPI_Aint *address);
int MPI_Address(void *location, MPI_Aint *address);
int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op);
int MPI_Op_free(MPI_Op *op);

#ifdef __cplusplus
}
#endif
#endif
