// oracle/emu/nccl_emu.cpp — TEST INFRASTRUCTURE: the NCCL calls of libhb200 over oracle/minimpi, so that
// the multi-rank path (halo exchange, all-reduced dots, coarse gather) can run as N host processes.
// Point-to-point calls inside a group are queued and run at ncclGroupEnd: all sends first (minimpi
// sends are eager and buffered), then the receives — the order-free completion NCCL groups give.
#include "nccl.h"
#include "../minimpi/mpi.h"
#include <stdlib.h>
#include <string.h>
#include <vector>

struct ncclComm { MPI_Comm comm; int rank, size; };

namespace {
struct P2P { bool send; void *buf; size_t bytes; int peer; MPI_Comm comm; };
int g_group = 0;
std::vector<P2P> g_queue;
const int kTag = 4242;

size_t dt_size(ncclDataType_t dt)
{
   switch (dt) {
      case ncclInt8: case ncclUint8: return 1;
      case ncclFloat16: return 2;
      case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
      default: return 8;
   }
}

void run(const P2P &op)
{
   if (op.send) MPI_Send(op.buf, (int) op.bytes, MPI_BYTE, op.peer, kTag, op.comm);
   else         MPI_Recv(op.buf, (int) op.bytes, MPI_BYTE, op.peer, kTag, op.comm, MPI_STATUS_IGNORE);
}

void ensure_mpi()
{
   int on = 0;
   MPI_Initialized(&on);
   if (!on) MPI_Init(nullptr, nullptr);
}
}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) { memset(id, 0, sizeof(*id)); strcpy(id->internal, "hb_emu"); return ncclSuccess; }

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId, int rank)
{
   ensure_mpi();
   ncclComm *c = new ncclComm();
   MPI_Comm_dup(MPI_COMM_WORLD, &c->comm);
   MPI_Comm_rank(c->comm, &c->rank);
   MPI_Comm_size(c->comm, &c->size);
   if (c->rank != rank || c->size != nranks) { delete c; return ncclInvalidArgument; }
   *comm = c;
   return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) { delete comm; return ncclSuccess; }

ncclResult_t ncclAllReduce(const void *sendbuff, void *recvbuff, size_t count, ncclDataType_t dt, ncclRedOp_t op,
                           ncclComm_t comm, hb_emu_stream_t)
{
   if (dt != ncclDouble || op != ncclSum) return ncclInvalidArgument;
   std::vector<double> tmp((const double *) sendbuff, (const double *) sendbuff + count);
   MPI_Allreduce(tmp.data(), recvbuff, (int) count, MPI_DOUBLE, MPI_SUM, comm->comm);
   return ncclSuccess;
}

ncclResult_t ncclAllGather(const void *sendbuff, void *recvbuff, size_t sendcount, ncclDataType_t dt, ncclComm_t comm,
                           hb_emu_stream_t)
{
   const size_t bytes = sendcount * dt_size(dt);
   std::vector<char> tmp((const char *) sendbuff, (const char *) sendbuff + bytes);   // sendbuff may lie inside recvbuff
   MPI_Allgather(tmp.data(), (int) bytes, MPI_BYTE, recvbuff, (int) bytes, MPI_BYTE, comm->comm);
   return ncclSuccess;
}

ncclResult_t ncclSend(const void *sendbuff, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, hb_emu_stream_t)
{
   P2P op{true, (void *) sendbuff, count * dt_size(dt), peer, comm->comm};
   if (g_group > 0) g_queue.push_back(op); else run(op);
   return ncclSuccess;
}

ncclResult_t ncclRecv(void *recvbuff, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, hb_emu_stream_t)
{
   P2P op{false, recvbuff, count * dt_size(dt), peer, comm->comm};
   if (g_group > 0) g_queue.push_back(op); else run(op);
   return ncclSuccess;
}

ncclResult_t ncclGroupStart(void) { g_group++; return ncclSuccess; }

ncclResult_t ncclGroupEnd(void)
{
   if (--g_group > 0) return ncclSuccess;
   for (const P2P &op : g_queue) if (op.send) run(op);
   for (const P2P &op : g_queue) if (!op.send) run(op);
   g_queue.clear();
   return ncclSuccess;
}

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }

}  // extern "C"
