/*
 * minimpi.c — implementation of the MPI subset declared in mpi.h (see that header for scope).
 *
 * Design: every rank owns a mailbox (multi-producer / single-consumer ring of fixed-size
 * envelopes, process-shared mutex + condvar) in one POSIX shared-memory segment per job.
 * Sends are eager and buffered: payloads <= MM_INLINE bytes travel inside the envelope, larger
 * ones in a per-message tmpfs file named in the envelope.  The receiver keeps the usual two
 * queues (posted receives in posting order, unexpected messages in arrival order), which gives
 * MPI's non-overtaking matching semantics including MPI_ANY_SOURCE / MPI_ANY_TAG and probes.
 * Collectives are built on point-to-point in a shadow context with rank-ordered (therefore
 * deterministic) reductions.  Single-threaded callers only (hypre calls MPI outside OpenMP
 * regions).
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#define MM_MAGIC   0x6d6d7069u
#define MM_NSLOT   4096
#define MM_INLINE  224
#define MM_MAXRANK 64

typedef struct
{
   int                src;      /* world rank */
   int                tag;
   int                ctx;
   int                has_file;
   long long          nbytes;
   unsigned long long seq;
   char               data[MM_INLINE];
} mm_slot;

typedef struct
{
   pthread_mutex_t    mu;
   pthread_cond_t     cv;
   unsigned long long head, tail;
   mm_slot            slots[MM_NSLOT];
} mm_box;

typedef struct
{
   volatile unsigned  magic;
   int                nranks;
   volatile int       go;
   volatile int       attached[MM_MAXRANK];
   volatile int       finalized[MM_MAXRANK];
   long long          create_ns;
   char               payload_dir[96];   /* where large payload files go (tmpfs if it is big enough) */
   mm_box             box[];
} mm_shared;

/* ---- private state ---------------------------------------------------------------------- */

typedef struct msg
{
   struct msg        *next;
   int                src, tag, ctx, has_file;
   long long          nbytes;
   unsigned long long seq;
   char              *data;
} msg;

typedef struct
{
   int          used, kind /* 1 send, 2 recv */, persistent, active, done;
   void        *buf;
   int          count;
   MPI_Datatype dt;
   int          peer;   /* comm rank or MPI_ANY_SOURCE */
   int          tag;
   MPI_Comm     comm;
   MPI_Status   st;
   int          next_posted;
} mm_req;

typedef struct
{
   int  used, ctx, size, rank;
   int *ranks;   /* world ranks */
} mm_comm;

typedef struct
{
   int  used, size;
   int *ranks;   /* world ranks */
} mm_group;

typedef struct
{
   int        used, builtin, contiguous, nblocks;
   long long  size, extent;
   long long *off, *len;
} mm_dtype;

typedef struct
{
   int                used;
   MPI_User_function *fn;
} mm_op;

static mm_shared *G = NULL;
static size_t     g_shm_bytes = 0;
static int        g_rank = 0, g_size = 1, g_init = 0, g_final = 0;
static char       g_job[160];
static unsigned long long g_seq = 0;

static msg *uq_head = NULL, *uq_tail = NULL;
static mm_req *reqs = NULL;
static int nreqs = 0;
static int posted_head = 0, posted_tail = 0;

static mm_comm  *comms = NULL;  static int ncomms = 0;
static mm_group *groups = NULL; static int ngroups = 0;
static mm_dtype *dts = NULL;    static int ndts = 0;
static mm_op    *ops = NULL;    static int nops = 0;
static int next_ctx = 16;

static void die(const char *what)
{
   fprintf(stderr, "[minimpi rank %d] fatal: %s (errno %d: %s)\n", g_rank, what, errno, strerror(errno));
   fflush(stderr);
   _exit(86);
}

static long long now_ns(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_REALTIME, &ts);
   return (long long) ts.tv_sec * 1000000000LL + ts.tv_nsec;
}

/* ---- datatypes -------------------------------------------------------------------------- */

static int dt_new(void)
{
   int i;
   for (i = 32; i < ndts; i++) if (!dts[i].used) { memset(&dts[i], 0, sizeof(mm_dtype)); dts[i].used = 1; return i; }
   {
      int old = ndts;
      ndts = ndts ? 2 * ndts : 64;
      dts = (mm_dtype *) realloc(dts, sizeof(mm_dtype) * ndts);
      memset(dts + old, 0, sizeof(mm_dtype) * (ndts - old));
      i = old < 32 ? 32 : old;
      dts[i].used = 1;
      return i;
   }
}

static void dt_builtin(int id, long long size)
{
   dts[id].used = 1; dts[id].builtin = 1; dts[id].contiguous = 1; dts[id].size = size; dts[id].extent = size;
   dts[id].nblocks = 1;
   dts[id].off = (long long *) malloc(sizeof(long long)); dts[id].len = (long long *) malloc(sizeof(long long));
   dts[id].off[0] = 0; dts[id].len[0] = size;
}

static void dt_init(void)
{
   ndts = 64;
   dts = (mm_dtype *) calloc(ndts, sizeof(mm_dtype));
   dt_builtin(MPI_CHAR, 1); dt_builtin(MPI_BYTE, 1); dt_builtin(MPI_SHORT, sizeof(short));
   dt_builtin(MPI_INT, sizeof(int)); dt_builtin(MPI_LONG, sizeof(long));
   dt_builtin(MPI_LONG_LONG_INT, sizeof(long long)); dt_builtin(MPI_UNSIGNED, sizeof(unsigned));
   dt_builtin(MPI_FLOAT, sizeof(float)); dt_builtin(MPI_DOUBLE, sizeof(double));
   dt_builtin(MPI_LONG_DOUBLE, sizeof(long double));
   dt_builtin(MPI_C_FLOAT_COMPLEX, 2 * sizeof(float)); dt_builtin(MPI_C_DOUBLE_COMPLEX, 2 * sizeof(double));
   dt_builtin(MPI_C_LONG_DOUBLE_COMPLEX, 2 * sizeof(long double));
   dt_builtin(MPI_UNSIGNED_LONG, sizeof(unsigned long));
}

static mm_dtype *dt_get(MPI_Datatype d)
{
   if (d <= 0 || d >= ndts || !dts[d].used) { fprintf(stderr, "[minimpi] bad datatype %d\n", d); die("datatype"); }
   return &dts[d];
}

static void dt_push(mm_dtype *t, long long off, long long len)
{
   if (len <= 0) return;
   if (t->nblocks > 0 && t->off[t->nblocks - 1] + t->len[t->nblocks - 1] == off) { t->len[t->nblocks - 1] += len; return; }
   t->off = (long long *) realloc(t->off, sizeof(long long) * (t->nblocks + 1));
   t->len = (long long *) realloc(t->len, sizeof(long long) * (t->nblocks + 1));
   t->off[t->nblocks] = off; t->len[t->nblocks] = len; t->nblocks++;
}

static void dt_append(mm_dtype *t, const mm_dtype *o, long long shift)
{
   int b;
   for (b = 0; b < o->nblocks; b++) dt_push(t, shift + o->off[b], o->len[b]);
   t->size += o->size;
}

static void dt_finish(mm_dtype *t, long long extent)
{
   t->extent = extent;
   t->contiguous = (t->nblocks == 1 && t->off[0] == 0 && t->len[0] == t->size && t->extent == t->size) || t->size == 0;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype)
{
   int id = dt_new(), i;
   mm_dtype *o = dt_get(oldtype), *t = &dts[id];
   for (i = 0; i < count; i++) dt_append(t, o, (long long) i * o->extent);
   dt_finish(t, (long long) count * o->extent);
   *newtype = id;
   return MPI_SUCCESS;
}

int MPI_Type_create_hvector(int count, int bl, MPI_Aint stride, MPI_Datatype oldtype, MPI_Datatype *newtype)
{
   int id = dt_new(), i, j;
   mm_dtype *o = dt_get(oldtype), *t = &dts[id];
   for (i = 0; i < count; i++) for (j = 0; j < bl; j++) dt_append(t, o, (long long) i * stride + (long long) j * o->extent);
   dt_finish(t, count > 0 ? (long long) (count - 1) * stride + (long long) bl * o->extent : 0);
   *newtype = id;
   return MPI_SUCCESS;
}
int MPI_Type_hvector(int count, int bl, MPI_Aint stride, MPI_Datatype o, MPI_Datatype *n) { return MPI_Type_create_hvector(count, bl, stride, o, n); }

int MPI_Type_vector(int count, int bl, int stride, MPI_Datatype oldtype, MPI_Datatype *newtype)
{
   mm_dtype *o = dt_get(oldtype);
   return MPI_Type_create_hvector(count, bl, (MPI_Aint) stride * o->extent, oldtype, newtype);
}

int MPI_Type_create_struct(int count, const int *bls, const MPI_Aint *displs, const MPI_Datatype *types, MPI_Datatype *newtype)
{
   int id = dt_new(), k, j;
   mm_dtype *t = &dts[id];
   long long ub = 0;
   for (k = 0; k < count; k++)
   {
      mm_dtype *o = dt_get(types[k]);
      for (j = 0; j < bls[k]; j++) dt_append(t, o, (long long) displs[k] + (long long) j * o->extent);
      if ((long long) displs[k] + (long long) bls[k] * o->extent > ub) ub = (long long) displs[k] + (long long) bls[k] * o->extent;
   }
   dt_finish(t, ub);
   t->contiguous = 0;   /* displacements may be absolute addresses (MPI_BOTTOM) */
   if (t->nblocks == 1 && t->off[0] == 0 && t->len[0] == t->size) t->contiguous = 1;
   *newtype = id;
   return MPI_SUCCESS;
}
int MPI_Type_struct(int count, int *bls, MPI_Aint *displs, MPI_Datatype *types, MPI_Datatype *n) { return MPI_Type_create_struct(count, bls, displs, types, n); }
int MPI_Type_commit(MPI_Datatype *dt) { (void) dt; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *dt)
{
   if (*dt >= 32 && *dt < ndts && dts[*dt].used) { free(dts[*dt].off); free(dts[*dt].len); memset(&dts[*dt], 0, sizeof(mm_dtype)); }
   *dt = MPI_DATATYPE_NULL;
   return MPI_SUCCESS;
}
int MPI_Type_size(MPI_Datatype dt, int *size) { *size = (int) dt_get(dt)->size; return MPI_SUCCESS; }
int MPI_Get_address(const void *location, MPI_Aint *address) { *address = (MPI_Aint) (size_t) location; return MPI_SUCCESS; }
int MPI_Address(void *location, MPI_Aint *address) { return MPI_Get_address(location, address); }

static void dt_pack(const void *buf, int count, mm_dtype *t, char *out)
{
   int e, b;
   if (t->contiguous) { memcpy(out, buf, (size_t) (t->size * count)); return; }
   for (e = 0; e < count; e++)
      for (b = 0; b < t->nblocks; b++)
      {
         memcpy(out, (const char *) buf + (long long) e * t->extent + t->off[b], (size_t) t->len[b]);
         out += t->len[b];
      }
}

static void dt_unpack(void *buf, int count, mm_dtype *t, const char *in, long long nbytes)
{
   int e, b;
   long long cap = t->size * count;
   if (nbytes > cap) nbytes = cap;
   if (t->contiguous) { memcpy(buf, in, (size_t) nbytes); return; }
   for (e = 0; e < count && nbytes > 0; e++)
      for (b = 0; b < t->nblocks && nbytes > 0; b++)
      {
         long long l = t->len[b] < nbytes ? t->len[b] : nbytes;
         memcpy((char *) buf + (long long) e * t->extent + t->off[b], in, (size_t) l);
         in += l; nbytes -= l;
      }
}

/* ---- communicators / groups ----------------------------------------------------------------- */

static int comm_new(void)
{
   int i;
   for (i = 3; i < ncomms; i++) if (!comms[i].used) return i;
   {
      int old = ncomms;
      ncomms = ncomms ? 2 * ncomms : 16;
      comms = (mm_comm *) realloc(comms, sizeof(mm_comm) * ncomms);
      memset(comms + old, 0, sizeof(mm_comm) * (ncomms - old));
      return old < 3 ? 3 : old;
   }
}

static mm_comm *comm_get(MPI_Comm c)
{
   if (c <= 0 || c >= ncomms || !comms[c].used) { fprintf(stderr, "[minimpi rank %d] bad communicator %d\n", g_rank, c); die("communicator"); }
   return &comms[c];
}

static int comm_rank_of_world(mm_comm *c, int world)
{
   int i;
   for (i = 0; i < c->size; i++) if (c->ranks[i] == world) return i;
   return MPI_UNDEFINED;
}

static MPI_Comm comm_make(int ctx, int size, const int *ranks)
{
   int id = comm_new(), i;
   comms[id].used = 1; comms[id].ctx = ctx; comms[id].size = size;
   comms[id].ranks = (int *) malloc(sizeof(int) * (size ? size : 1));
   comms[id].rank = MPI_UNDEFINED;
   for (i = 0; i < size; i++) { comms[id].ranks[i] = ranks[i]; if (ranks[i] == g_rank) comms[id].rank = i; }
   return id;
}

int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm_get(comm)->size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm_get(comm)->rank; return MPI_SUCCESS; }
MPI_Comm MPI_Comm_f2c(MPI_Fint comm) { return (MPI_Comm) comm; }
MPI_Fint MPI_Comm_c2f(MPI_Comm comm) { return (MPI_Fint) comm; }
int MPI_Info_create(MPI_Info *info) { *info = 1; return MPI_SUCCESS; }
int MPI_Info_free(MPI_Info *info) { *info = MPI_INFO_NULL; return MPI_SUCCESS; }

static int group_new(void)
{
   int i;
   for (i = 1; i < ngroups; i++) if (!groups[i].used) return i;
   {
      int old = ngroups;
      ngroups = ngroups ? 2 * ngroups : 16;
      groups = (mm_group *) realloc(groups, sizeof(mm_group) * ngroups);
      memset(groups + old, 0, sizeof(mm_group) * (ngroups - old));
      return old < 1 ? 1 : old;
   }
}

int MPI_Comm_group(MPI_Comm comm, MPI_Group *group)
{
   mm_comm *c = comm_get(comm);
   int g = group_new();
   groups[g].used = 1; groups[g].size = c->size;
   groups[g].ranks = (int *) malloc(sizeof(int) * (c->size ? c->size : 1));
   memcpy(groups[g].ranks, c->ranks, sizeof(int) * c->size);
   *group = g;
   return MPI_SUCCESS;
}

int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup)
{
   int g = group_new(), i;
   groups[g].used = 1; groups[g].size = n;
   groups[g].ranks = (int *) malloc(sizeof(int) * (n ? n : 1));
   for (i = 0; i < n; i++) groups[g].ranks[i] = groups[group].ranks[ranks[i]];
   *newgroup = g;
   return MPI_SUCCESS;
}

int MPI_Group_free(MPI_Group *group)
{
   if (*group > 0 && *group < ngroups && groups[*group].used) { free(groups[*group].ranks); groups[*group].used = 0; }
   *group = MPI_GROUP_NULL;
   return MPI_SUCCESS;
}

/* ---- mailbox ---------------------------------------------------------------------------------- */

static void payload_path(char *out, size_t n, int src, unsigned long long seq)
{
   snprintf(out, n, "%s/minimpi_%s_m_%d_%llu", G->payload_dir, g_job, src, seq);
}

static int req_matches(const mm_req *r, const msg *m)
{
   mm_comm *c = &comms[r->comm];
   if (c->ctx + (r->kind == 3 ? 1 : 0) != m->ctx) return 0;
   if (r->tag != MPI_ANY_TAG && r->tag != m->tag) return 0;
   if (r->peer != MPI_ANY_SOURCE && c->ranks[r->peer] != m->src) return 0;
   return 1;
}

static void read_payload(const msg *m, void *buf, int count, MPI_Datatype dt)
{
   mm_dtype *t = dt_get(dt);
   long long cap = t->size * count;
   long long n = m->nbytes < cap ? m->nbytes : cap;
   if (m->nbytes > cap)
   {
      fprintf(stderr, "[minimpi rank %d] message truncated: %lld bytes into %lld (src %d tag %d)\n",
              g_rank, m->nbytes, cap, m->src, m->tag);
   }
   if (!m->has_file)
   {
      dt_unpack(buf, count, t, m->data, n);
      return;
   }
   {
      char path[256];
      int fd;
      long long got = 0;
      char *tmp = NULL, *dst;
      payload_path(path, sizeof(path), m->src, m->seq);
      fd = open(path, O_RDONLY);
      if (fd < 0) die("open payload");
      if (t->contiguous) dst = (char *) buf; else { tmp = (char *) malloc((size_t) (n ? n : 1)); dst = tmp; }
      while (got < n)
      {
         ssize_t k = read(fd, dst + got, (size_t) (n - got));
         if (k < 0) { if (errno == EINTR) continue; die("read payload"); }
         if (k == 0) break;
         got += k;
      }
      close(fd);
      unlink(path);
      if (tmp) { dt_unpack(buf, count, t, tmp, n); free(tmp); }
   }
}

static void complete_recv(mm_req *r, msg *m)
{
   mm_comm *c = &comms[r->comm];
   read_payload(m, r->buf, r->count, r->dt);
   r->st.MPI_SOURCE = comm_rank_of_world(c, m->src);
   r->st.MPI_TAG = m->tag;
   r->st.MPI_ERROR = MPI_SUCCESS;
   r->st._nbytes = m->nbytes;
   r->done = 1;
}

static void posted_remove(int idx)
{
   int p = posted_head, prev = 0;
   while (p)
   {
      if (p == idx)
      {
         if (prev) reqs[prev].next_posted = reqs[p].next_posted; else posted_head = reqs[p].next_posted;
         if (posted_tail == p) posted_tail = prev;
         reqs[p].next_posted = 0;
         return;
      }
      prev = p; p = reqs[p].next_posted;
   }
}

/* drain this rank's mailbox: match against posted receives, else queue as unexpected */
static void progress(void)
{
   mm_box *b = &G->box[g_rank];
   for (;;)
   {
      mm_slot s;
      msg *m;
      int p;
      pthread_mutex_lock(&b->mu);
      if (b->head == b->tail) { pthread_mutex_unlock(&b->mu); return; }
      s = b->slots[b->head % MM_NSLOT];
      b->head++;
      pthread_mutex_unlock(&b->mu);

      m = (msg *) malloc(sizeof(msg));
      m->next = NULL; m->src = s.src; m->tag = s.tag; m->ctx = s.ctx; m->has_file = s.has_file;
      m->nbytes = s.nbytes; m->seq = s.seq; m->data = NULL;
      if (!s.has_file && s.nbytes > 0) { m->data = (char *) malloc((size_t) s.nbytes); memcpy(m->data, s.data, (size_t) s.nbytes); }
      for (p = posted_head; p; p = reqs[p].next_posted)
      {
         if (req_matches(&reqs[p], m))
         {
            complete_recv(&reqs[p], m);
            posted_remove(p);
            free(m->data); free(m);
            m = NULL;
            break;
         }
      }
      if (m) { if (uq_tail) uq_tail->next = m; else uq_head = m; uq_tail = m; }
   }
}

static void wait_for_mail(void)
{
   mm_box *b = &G->box[g_rank];
   struct timespec ts;
   pthread_mutex_lock(&b->mu);
   if (b->head == b->tail)
   {
      clock_gettime(CLOCK_REALTIME, &ts);
      ts.tv_nsec += 1000000;   /* 1 ms */
      if (ts.tv_nsec >= 1000000000L) { ts.tv_sec++; ts.tv_nsec -= 1000000000L; }
      pthread_cond_timedwait(&b->cv, &b->mu, &ts);
   }
   pthread_mutex_unlock(&b->mu);
}

static void raw_send(const char *bytes, long long nbytes, int dest_world, int tag, int ctx)
{
   mm_box *b = &G->box[dest_world];
   mm_slot s;
   s.src = g_rank; s.tag = tag; s.ctx = ctx; s.nbytes = nbytes; s.has_file = 0; s.seq = 0;
   if (nbytes <= MM_INLINE) { if (nbytes > 0) memcpy(s.data, bytes, (size_t) nbytes); }
   else
   {
      char path[256];
      int fd;
      long long put = 0;
      s.has_file = 1; s.seq = ++g_seq;
      payload_path(path, sizeof(path), g_rank, s.seq);
      fd = open(path, O_CREAT | O_EXCL | O_WRONLY, 0600);
      if (fd < 0) die("create payload");
      while (put < nbytes)
      {
         ssize_t k = write(fd, bytes + put, (size_t) (nbytes - put));
         if (k < 0) { if (errno == EINTR) continue; die("write payload (is /dev/shm full?)"); }
         put += k;
      }
      close(fd);
   }
   for (;;)
   {
      pthread_mutex_lock(&b->mu);
      if (b->tail - b->head < MM_NSLOT)
      {
         b->slots[b->tail % MM_NSLOT] = s;
         b->tail++;
         pthread_cond_signal(&b->cv);
         pthread_mutex_unlock(&b->mu);
         return;
      }
      pthread_mutex_unlock(&b->mu);
      progress();     /* the peer may be blocked sending to us */
      usleep(50);
   }
}

static void typed_send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, int coll)
{
   mm_comm *c = comm_get(comm);
   mm_dtype *t = dt_get(dt);
   long long nbytes = t->size * count;
   if (dest == MPI_PROC_NULL) return;
   if (dest < 0 || dest >= c->size) { fprintf(stderr, "[minimpi rank %d] send to bad rank %d\n", g_rank, dest); die("send"); }
   if (t->contiguous) raw_send((const char *) buf, nbytes, c->ranks[dest], tag, c->ctx + coll);
   else
   {
      char *tmp = (char *) malloc((size_t) (nbytes ? nbytes : 1));
      dt_pack(buf, count, t, tmp);
      raw_send(tmp, nbytes, c->ranks[dest], tag, c->ctx + coll);
      free(tmp);
   }
}

/* ---- requests ------------------------------------------------------------------------------------ */

static int req_new(void)
{
   int i;
   for (i = 1; i < nreqs; i++) if (!reqs[i].used) { memset(&reqs[i], 0, sizeof(mm_req)); reqs[i].used = 1; return i; }
   {
      int old = nreqs;
      nreqs = nreqs ? 2 * nreqs : 256;
      reqs = (mm_req *) realloc(reqs, sizeof(mm_req) * nreqs);
      memset(reqs + old, 0, sizeof(mm_req) * (nreqs - old));
      i = old < 1 ? 1 : old;
      reqs[i].used = 1;
      return i;
   }
}

/* post a receive: try the unexpected queue first (arrival order), else append to posted */
static void post_recv(int idx)
{
   mm_req *r = &reqs[idx];
   msg *m, *prev = NULL;
   progress();
   r = &reqs[idx];
   for (m = uq_head; m; prev = m, m = m->next)
   {
      if (req_matches(r, m))
      {
         if (prev) prev->next = m->next; else uq_head = m->next;
         if (uq_tail == m) uq_tail = prev;
         complete_recv(r, m);
         free(m->data); free(m);
         return;
      }
   }
   r->next_posted = 0;
   if (posted_tail) reqs[posted_tail].next_posted = idx; else posted_head = idx;
   posted_tail = idx;
}

static void wait_req(int idx)
{
   while (!reqs[idx].done) { progress(); if (reqs[idx].done) break; wait_for_mail(); }
}

static void finish_req(MPI_Request *req, MPI_Status *status)
{
   mm_req *r = &reqs[*req];
   if (status && r->kind >= 2) *status = r->st;
   if (r->persistent) { r->active = 0; r->done = 0; }
   else { r->used = 0; *req = MPI_REQUEST_NULL; }
}

int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
   int idx;
   typed_send(buf, count, dt, dest, tag, comm, 0);
   idx = req_new();
   reqs[idx].kind = 1; reqs[idx].done = 1; reqs[idx].active = 1;
   *req = idx;
   return MPI_SUCCESS;
}
int MPI_Irsend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
   return MPI_Isend(buf, count, dt, dest, tag, comm, req);
}
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm)
{
   typed_send(buf, count, dt, dest, tag, comm, 0);
   return MPI_SUCCESS;
}

static int irecv_kind(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, int kind)
{
   int idx = req_new();
   mm_req *r = &reqs[idx];
   r->kind = kind; r->buf = buf; r->count = count; r->dt = dt; r->peer = source; r->tag = tag; r->comm = comm;
   r->active = 1;
   if (source == MPI_PROC_NULL)
   {
      r->done = 1; r->st.MPI_SOURCE = MPI_PROC_NULL; r->st.MPI_TAG = MPI_ANY_TAG; r->st._nbytes = 0;
      return idx;
   }
   post_recv(idx);
   return idx;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req)
{
   comm_get(comm);
   *req = irecv_kind(buf, count, dt, source, tag, comm, 2);
   return MPI_SUCCESS;
}

int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status)
{
   MPI_Request r;
   MPI_Irecv(buf, count, dt, source, tag, comm, &r);
   return MPI_Wait(&r, status);
}

int MPI_Send_init(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
   int idx = req_new();
   mm_req *r = &reqs[idx];
   r->kind = 1; r->persistent = 1; r->buf = (void *) buf; r->count = count; r->dt = dt; r->peer = dest; r->tag = tag; r->comm = comm;
   *req = idx;
   return MPI_SUCCESS;
}
int MPI_Recv_init(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req)
{
   int idx = req_new();
   mm_req *r = &reqs[idx];
   r->kind = 2; r->persistent = 1; r->buf = buf; r->count = count; r->dt = dt; r->peer = source; r->tag = tag; r->comm = comm;
   *req = idx;
   return MPI_SUCCESS;
}
int MPI_Start(MPI_Request *req)
{
   mm_req *r = &reqs[*req];
   r->active = 1; r->done = 0;
   if (r->kind == 1) { typed_send(r->buf, r->count, r->dt, r->peer, r->tag, r->comm, 0); reqs[*req].done = 1; }
   else post_recv(*req);
   return MPI_SUCCESS;
}
int MPI_Startall(int count, MPI_Request *rq) { int i; for (i = 0; i < count; i++) MPI_Start(&rq[i]); return MPI_SUCCESS; }

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{
   if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
   if (reqs[*req].persistent && !reqs[*req].active) return MPI_SUCCESS;
   wait_req(*req);
   finish_req(req, status);
   return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request *rq, MPI_Status *statuses)
{
   int i;
   for (i = 0; i < count; i++) MPI_Wait(&rq[i], statuses ? &statuses[i] : NULL);
   return MPI_SUCCESS;
}
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status)
{
   *flag = 1;
   if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
   if (reqs[*req].persistent && !reqs[*req].active) return MPI_SUCCESS;
   progress();
   if (reqs[*req].done) finish_req(req, status); else *flag = 0;
   return MPI_SUCCESS;
}
int MPI_Testall(int count, MPI_Request *rq, int *flag, MPI_Status *statuses)
{
   int i;
   progress();
   for (i = 0; i < count; i++)
   {
      if (rq[i] == MPI_REQUEST_NULL) continue;
      if (reqs[rq[i]].persistent && !reqs[rq[i]].active) continue;
      if (!reqs[rq[i]].done) { *flag = 0; return MPI_SUCCESS; }
   }
   for (i = 0; i < count; i++) if (rq[i] != MPI_REQUEST_NULL) finish_req(&rq[i], statuses ? &statuses[i] : NULL);
   *flag = 1;
   return MPI_SUCCESS;
}
int MPI_Waitany(int count, MPI_Request *rq, int *index, MPI_Status *status)
{
   int i, any;
   for (;;)
   {
      any = 0;
      progress();
      for (i = 0; i < count; i++)
      {
         if (rq[i] == MPI_REQUEST_NULL) continue;
         if (reqs[rq[i]].persistent && !reqs[rq[i]].active) continue;
         any = 1;
         if (reqs[rq[i]].done) { finish_req(&rq[i], status); *index = i; return MPI_SUCCESS; }
      }
      if (!any) { *index = MPI_UNDEFINED; return MPI_SUCCESS; }
      wait_for_mail();
   }
}
int MPI_Request_free(MPI_Request *req)
{
   if (*req != MPI_REQUEST_NULL)
   {
      if (reqs[*req].kind >= 2 && !reqs[*req].done) posted_remove(*req);
      reqs[*req].used = 0;
      *req = MPI_REQUEST_NULL;
   }
   return MPI_SUCCESS;
}

static int probe_once(int source, int tag, MPI_Comm comm, MPI_Status *status)
{
   mm_comm *c = comm_get(comm);
   mm_req r;
   msg *m;
   memset(&r, 0, sizeof(r));
   r.kind = 2; r.peer = source; r.tag = tag; r.comm = comm;
   progress();
   for (m = uq_head; m; m = m->next)
   {
      if (req_matches(&r, m))
      {
         if (status) { status->MPI_SOURCE = comm_rank_of_world(c, m->src); status->MPI_TAG = m->tag; status->MPI_ERROR = 0; status->_nbytes = m->nbytes; }
         return 1;
      }
   }
   return 0;
}
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int *flag, MPI_Status *status)
{
   *flag = probe_once(source, tag, comm, status);
   return MPI_SUCCESS;
}
int MPI_Probe(int source, int tag, MPI_Comm comm, MPI_Status *status)
{
   while (!probe_once(source, tag, comm, status)) wait_for_mail();
   return MPI_SUCCESS;
}
int MPI_Get_count(const MPI_Status *status, MPI_Datatype dt, int *count)
{
   long long sz = dt_get(dt)->size;
   *count = sz ? (int) (status->_nbytes / sz) : 0;
   return MPI_SUCCESS;
}

/* ---- collectives (shadow context ctx+1) ----------------------------------------------------------- */

static void csend(MPI_Comm comm, int dest, int tag, const void *buf, long long nbytes)
{
   mm_comm *c = comm_get(comm);
   raw_send((const char *) buf, nbytes, c->ranks[dest], tag, c->ctx + 1);
}
static void crecv(MPI_Comm comm, int src, int tag, void *buf, long long nbytes)
{
   int idx = irecv_kind(buf, (int) nbytes, MPI_BYTE, src, tag, comm, 3);
   MPI_Request r = idx;
   wait_req(idx);
   finish_req(&r, NULL);
}

enum { T_BARRIER = 1, T_BCAST, T_REDUCE, T_SCAN, T_GATHER, T_SCATTER, T_ALLTOALL, T_CTX };

int MPI_Barrier(MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   char tok = 0;
   int i;
   if (c->size == 1) return MPI_SUCCESS;
   if (c->rank == 0)
   {
      for (i = 1; i < c->size; i++) crecv(comm, i, T_BARRIER, &tok, 1);
      for (i = 1; i < c->size; i++) csend(comm, i, T_BARRIER, &tok, 1);
   }
   else { csend(comm, 0, T_BARRIER, &tok, 1); crecv(comm, 0, T_BARRIER, &tok, 1); }
   return MPI_SUCCESS;
}

static void bcast_bytes(void *buf, long long nbytes, int root, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   int i;
   if (c->size == 1) return;
   if (c->rank == root) { for (i = 0; i < c->size; i++) if (i != root) csend(comm, i, T_BCAST, buf, nbytes); }
   else crecv(comm, root, T_BCAST, buf, nbytes);
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   mm_dtype *t = dt_get(dt);
   long long nbytes = t->size * count;
   if (t->contiguous) bcast_bytes(buf, nbytes, root, comm);
   else
   {
      char *tmp = (char *) malloc((size_t) (nbytes ? nbytes : 1));
      if (c->rank == root) dt_pack(buf, count, t, tmp);
      bcast_bytes(tmp, nbytes, root, comm);
      if (c->rank != root) dt_unpack(buf, count, t, tmp, nbytes);
      free(tmp);
   }
   return MPI_SUCCESS;
}

#define RED_LOOP(T, EXPR) { T *a = (T *) inout; const T *b = (const T *) in; int i_; for (i_ = 0; i_ < n; i_++) { T x = a[i_], y = b[i_]; a[i_] = (EXPR); } }
#define RED_ARITH(T) switch (op) { \
   case MPI_SUM:  RED_LOOP(T, x + y) break; case MPI_PROD: RED_LOOP(T, x * y) break; \
   case MPI_MIN:  RED_LOOP(T, x < y ? x : y) break; case MPI_MAX: RED_LOOP(T, x > y ? x : y) break; \
   case MPI_LOR:  RED_LOOP(T, (x != 0) || (y != 0)) break; case MPI_LAND: RED_LOOP(T, (x != 0) && (y != 0)) break; \
   default: die("reduction op not supported for this datatype"); }
#define RED_INT(T) switch (op) { \
   case MPI_BOR: RED_LOOP(T, x | y) break; case MPI_BAND: RED_LOOP(T, x & y) break; default: RED_ARITH(T) }

/* inout = inout (op) in, element-wise; "in" is the contribution of the higher rank */
static void reduce_apply(MPI_Op op, MPI_Datatype dt, void *inout, const void *in, int n)
{
   if (op >= 16)
   {
      int len = n;
      MPI_Datatype d = dt;
      /* MPI: inoutvec[i] = invec[i] op inoutvec[i]; keep lower ranks on the left */
      char *tmp = (char *) malloc((size_t) (dt_get(dt)->size * n + 1));
      memcpy(tmp, in, (size_t) (dt_get(dt)->size * n));
      ops[op - 16].fn(inout, tmp, &len, &d);
      memcpy(inout, tmp, (size_t) (dt_get(dt)->size * n));
      free(tmp);
      return;
   }
   switch (dt)
   {
      case MPI_INT: RED_INT(int) break;
      case MPI_LONG: RED_INT(long) break;
      case MPI_LONG_LONG_INT: RED_INT(long long) break;
      case MPI_UNSIGNED: RED_INT(unsigned) break;
      case MPI_UNSIGNED_LONG: RED_INT(unsigned long) break;
      case MPI_SHORT: RED_INT(short) break;
      case MPI_CHAR: case MPI_BYTE: RED_INT(char) break;
      case MPI_FLOAT: RED_ARITH(float) break;
      case MPI_DOUBLE: RED_ARITH(double) break;
      case MPI_LONG_DOUBLE: RED_ARITH(long double) break;
      case MPI_C_DOUBLE_COMPLEX:
         if (op == MPI_SUM) { n *= 2; { RED_LOOP(double, x + y) } } else die("complex reduction");
         break;
      default: die("reduction on a derived datatype");
   }
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   long long nbytes = dt_get(dt)->size * count;
   int i;
   if (c->rank != root) { csend(comm, root, T_REDUCE, sendbuf, nbytes); return MPI_SUCCESS; }
   {
      /* fold in rank order 0,1,...,n-1: deterministic */
      char *acc = (char *) malloc((size_t) (nbytes ? nbytes : 1));
      char *tmp = (char *) malloc((size_t) (nbytes ? nbytes : 1));
      for (i = 0; i < c->size; i++)
      {
         const void *contrib;
         if (i == root) contrib = sendbuf; else { crecv(comm, i, T_REDUCE, tmp, nbytes); contrib = tmp; }
         if (i == 0) memcpy(acc, contrib, (size_t) nbytes); else reduce_apply(op, dt, acc, contrib, count);
      }
      memcpy(recvbuf, acc, (size_t) nbytes);
      free(acc); free(tmp);
   }
   return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm)
{
   long long nbytes = dt_get(dt)->size * count;
   char *tmp = (char *) malloc((size_t) (nbytes ? nbytes : 1));
   memcpy(tmp, sendbuf, (size_t) nbytes);   /* sendbuf may alias recvbuf in sloppy callers */
   MPI_Reduce(tmp, recvbuf, count, dt, op, 0, comm);
   bcast_bytes(recvbuf, nbytes, 0, comm);
   free(tmp);
   return MPI_SUCCESS;
}

int MPI_Scan(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   long long nbytes = dt_get(dt)->size * count;
   char *mine = (char *) malloc((size_t) (nbytes ? nbytes : 1));
   memcpy(mine, sendbuf, (size_t) nbytes);
   if (c->rank > 0)
   {
      crecv(comm, c->rank - 1, T_SCAN, recvbuf, nbytes);
      reduce_apply(op, dt, recvbuf, mine, count);
   }
   else memcpy(recvbuf, mine, (size_t) nbytes);
   if (c->rank + 1 < c->size) csend(comm, c->rank + 1, T_SCAN, recvbuf, nbytes);
   free(mine);
   return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, const int *recvcounts,
                const int *displs, MPI_Datatype recvtype, int root, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   mm_dtype *st = dt_get(sendtype);
   int i;
   if (c->rank != root)
   {
      typed_send(sendbuf, sendcount, sendtype, root, T_GATHER, comm, 1);
      return MPI_SUCCESS;
   }
   {
      mm_dtype *rt = dt_get(recvtype);
      for (i = 0; i < c->size; i++)
      {
         char *dst = (char *) recvbuf + (long long) displs[i] * rt->extent;
         if (i == root)
         {
            char *tmp = (char *) malloc((size_t) (st->size * sendcount + 1));
            dt_pack(sendbuf, sendcount, st, tmp);
            dt_unpack(dst, recvcounts[i], rt, tmp, st->size * sendcount);
            free(tmp);
         }
         else
         {
            int idx = irecv_kind(dst, recvcounts[i], recvtype, i, T_GATHER, comm, 3);
            MPI_Request r = idx;
            wait_req(idx);
            finish_req(&r, NULL);
         }
      }
   }
   return MPI_SUCCESS;
}

int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
               MPI_Datatype recvtype, int root, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   int *cnt = (int *) malloc(sizeof(int) * c->size), *dsp = (int *) malloc(sizeof(int) * c->size), i, r;
   for (i = 0; i < c->size; i++) { cnt[i] = recvcount; dsp[i] = i * recvcount; }
   r = MPI_Gatherv(sendbuf, sendcount, sendtype, recvbuf, cnt, dsp, recvtype, root, comm);
   free(cnt); free(dsp);
   return r;
}

int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, const int *recvcounts,
                   const int *displs, MPI_Datatype recvtype, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   mm_dtype *rt = dt_get(recvtype);
   int i;
   /* the caller's send buffer may live inside recvbuf */
   long long sb = dt_get(sendtype)->size * sendcount;
   char *tmp = (char *) malloc((size_t) (sb ? sb : 1));
   dt_pack(sendbuf, sendcount, dt_get(sendtype), tmp);
   MPI_Gatherv(tmp, (int) sb, MPI_BYTE, recvbuf, recvcounts, displs, recvtype, 0, comm);
   free(tmp);
   /* broadcast every segment (segments may be non-contiguous in recvbuf) */
   for (i = 0; i < c->size; i++)
   {
      char *seg = (char *) recvbuf + (long long) displs[i] * rt->extent;
      if (rt->contiguous) bcast_bytes(seg, rt->size * recvcounts[i], 0, comm);
      else MPI_Bcast(seg, recvcounts[i], recvtype, 0, comm);
   }
   return MPI_SUCCESS;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   int *cnt = (int *) malloc(sizeof(int) * c->size), *dsp = (int *) malloc(sizeof(int) * c->size), i, r;
   for (i = 0; i < c->size; i++) { cnt[i] = recvcount; dsp[i] = i * recvcount; }
   r = MPI_Allgatherv(sendbuf, sendcount, sendtype, recvbuf, cnt, dsp, recvtype, comm);
   free(cnt); free(dsp);
   return r;
}

int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype sendtype, void *recvbuf,
                 int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   int i;
   if (c->rank == root)
   {
      mm_dtype *st = dt_get(sendtype);
      for (i = 0; i < c->size; i++)
      {
         const char *src = (const char *) sendbuf + (long long) displs[i] * st->extent;
         if (i == root)
         {
            char *tmp = (char *) malloc((size_t) (st->size * sendcounts[i] + 1));
            dt_pack(src, sendcounts[i], st, tmp);
            dt_unpack(recvbuf, recvcount, dt_get(recvtype), tmp, st->size * sendcounts[i]);
            free(tmp);
         }
         else typed_send(src, sendcounts[i], sendtype, i, T_SCATTER, comm, 1);
      }
   }
   else
   {
      int idx = irecv_kind(recvbuf, recvcount, recvtype, root, T_SCATTER, comm, 3);
      MPI_Request r = idx;
      wait_req(idx);
      finish_req(&r, NULL);
   }
   return MPI_SUCCESS;
}

int MPI_Scatter(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                MPI_Datatype recvtype, int root, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   int *cnt = (int *) malloc(sizeof(int) * c->size), *dsp = (int *) malloc(sizeof(int) * c->size), i, r;
   for (i = 0; i < c->size; i++) { cnt[i] = sendcount; dsp[i] = i * sendcount; }
   r = MPI_Scatterv(sendbuf, cnt, dsp, sendtype, recvbuf, recvcount, recvtype, root, comm);
   free(cnt); free(dsp);
   return r;
}

int MPI_Alltoall(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                 MPI_Datatype recvtype, MPI_Comm comm)
{
   mm_comm *c = comm_get(comm);
   mm_dtype *st = dt_get(sendtype), *rt = dt_get(recvtype);
   int i;
   for (i = 0; i < c->size; i++)
      typed_send((const char *) sendbuf + (long long) i * sendcount * st->extent, sendcount, sendtype, i, T_ALLTOALL, comm, 1);
   for (i = 0; i < c->size; i++)
   {
      int idx = irecv_kind((char *) recvbuf + (long long) i * recvcount * rt->extent, recvcount, recvtype, i, T_ALLTOALL, comm, 3);
      MPI_Request r = idx;
      wait_req(idx);
      finish_req(&r, NULL);
   }
   return MPI_SUCCESS;
}

/* ---- communicator construction (collective over the parent) ------------------------------------- */

static int agree_ctx(MPI_Comm parent)
{
   int mine = next_ctx, agreed = 0;
   MPI_Allreduce(&mine, &agreed, 1, MPI_INT, MPI_MAX, parent);
   next_ctx = agreed + 2;
   return agreed;
}

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm)
{
   mm_comm *c = comm_get(comm);
   int ctx = agree_ctx(comm);
   c = comm_get(comm);
   *newcomm = comm_make(ctx, c->size, c->ranks);
   return MPI_SUCCESS;
}

int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm)
{
   int ctx = agree_ctx(comm), i, member = 0;
   mm_group *g = (group > 0 && group < ngroups && groups[group].used) ? &groups[group] : NULL;
   if (g) for (i = 0; i < g->size; i++) if (g->ranks[i] == g_rank) member = 1;
   *newcomm = member ? comm_make(ctx, g->size, g->ranks) : MPI_COMM_NULL;
   return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm)
{
   mm_comm *c = comm_get(comm);
   int n = c->size, i, j, m = 0;
   int ctx = agree_ctx(comm);
   int mine[2], *all = (int *) malloc(sizeof(int) * 2 * n), *members = (int *) malloc(sizeof(int) * n);
   int *keys = (int *) malloc(sizeof(int) * n);
   c = comm_get(comm);
   mine[0] = color; mine[1] = key;
   MPI_Allgather(mine, 2, MPI_INT, all, 2, MPI_INT, comm);
   c = comm_get(comm);
   if (color == MPI_UNDEFINED) { *newcomm = MPI_COMM_NULL; free(all); free(members); free(keys); return MPI_SUCCESS; }
   for (i = 0; i < n; i++) if (all[2 * i] == color) { members[m] = c->ranks[i]; keys[m] = all[2 * i + 1]; m++; }
   /* stable insertion sort by key (ties keep parent-rank order) */
   for (i = 1; i < m; i++)
   {
      int kr = members[i], kk = keys[i];
      for (j = i - 1; j >= 0 && keys[j] > kk; j--) { members[j + 1] = members[j]; keys[j + 1] = keys[j]; }
      members[j + 1] = kr; keys[j + 1] = kk;
   }
   *newcomm = comm_make(ctx, m, members);
   free(all); free(members); free(keys);
   return MPI_SUCCESS;
}

int MPI_Comm_split_type(MPI_Comm comm, int split_type, int key, MPI_Info info, MPI_Comm *newcomm)
{
   (void) split_type; (void) info;   /* one node: everybody shares memory */
   return MPI_Comm_split(comm, 0, key, newcomm);
}

int MPI_Comm_free(MPI_Comm *comm)
{
   if (*comm > 2 && *comm < ncomms && comms[*comm].used) { free(comms[*comm].ranks); comms[*comm].used = 0; }
   *comm = MPI_COMM_NULL;
   return MPI_SUCCESS;
}

int MPI_Op_create(MPI_User_function *fn, int commute, MPI_Op *op)
{
   int i;
   (void) commute;
   for (i = 0; i < nops; i++) if (!ops[i].used) break;
   if (i == nops) { nops = nops ? 2 * nops : 8; ops = (mm_op *) realloc(ops, sizeof(mm_op) * nops); memset(ops + i, 0, sizeof(mm_op) * (nops - i)); }
   ops[i].used = 1; ops[i].fn = fn;
   *op = 16 + i;
   return MPI_SUCCESS;
}
int MPI_Op_free(MPI_Op *op) { if (*op >= 16) ops[*op - 16].used = 0; *op = MPI_OP_NULL; return MPI_SUCCESS; }

/* ---- init / finalize ---------------------------------------------------------------------------- */

static void cleanup_atexit(void)
{
   if (G && g_rank == 0 && g_init && !g_final)
   {
      char name[200];
      snprintf(name, sizeof(name), "/minimpi_%s", g_job);
      shm_unlink(name);
   }
}

int MPI_Initialized(int *flag) { *flag = g_init; return MPI_SUCCESS; }
int MPI_Finalized(int *flag) { *flag = g_final; return MPI_SUCCESS; }

int MPI_Init(int *argc, char ***argv)
{
   const char *e_rank, *e_size, *e_job;
   char name[200];
   int fd, i, world[MM_MAXRANK];
   (void) argc; (void) argv;
   if (g_init) return MPI_SUCCESS;
   e_rank = getenv("MINIMPI_RANK"); e_size = getenv("MINIMPI_SIZE"); e_job = getenv("MINIMPI_JOB");
   if (!e_rank || !e_size)
   {
      /* torchrun: RANK / WORLD_SIZE / MASTER_PORT */
      e_rank = getenv("RANK"); e_size = getenv("WORLD_SIZE");
      if (!e_job) e_job = getenv("MASTER_PORT");
   }
   g_rank = e_rank ? atoi(e_rank) : 0;
   g_size = e_size ? atoi(e_size) : 1;
   if (g_size < 1 || g_size > MM_MAXRANK || g_rank < 0 || g_rank >= g_size) die("bad MINIMPI_RANK / MINIMPI_SIZE");
   if (e_job && g_size > 1) snprintf(g_job, sizeof(g_job), "%s_%d", e_job, (int) getuid());
   else snprintf(g_job, sizeof(g_job), "solo%d", (int) getpid());
   snprintf(name, sizeof(name), "/minimpi_%s", g_job);
   g_shm_bytes = sizeof(mm_shared) + sizeof(mm_box) * (size_t) g_size;

   if (g_rank == 0)
   {
      pthread_mutexattr_t ma;
      pthread_condattr_t ca;
      shm_unlink(name);
      fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0) die("shm_open(create)");
      if (ftruncate(fd, (off_t) g_shm_bytes) != 0) die("ftruncate (is /dev/shm large enough?)");
      G = (mm_shared *) mmap(NULL, g_shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (G == MAP_FAILED) die("mmap");
      close(fd);
      memset((void *) G, 0, sizeof(mm_shared));
      pthread_mutexattr_init(&ma); pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
      pthread_condattr_init(&ca); pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
      for (i = 0; i < g_size; i++)
      {
         pthread_mutex_init(&G->box[i].mu, &ma);
         pthread_cond_init(&G->box[i].cv, &ca);
         G->box[i].head = G->box[i].tail = 0;
      }
      G->nranks = g_size;
      G->create_ns = now_ns();
      {
         /* large messages travel as files: tmpfs when it has room (containers often cap /dev/shm
            at 64 MB), else MINIMPI_TMPDIR or /tmp */
         const char *d = getenv("MINIMPI_TMPDIR");
         struct statvfs vf;
         if (!d)
         {
            d = "/tmp";
            if (statvfs("/dev/shm", &vf) == 0 && (double) vf.f_bavail * (double) vf.f_frsize > 8e9) { d = "/dev/shm"; }
         }
         snprintf((char *) G->payload_dir, sizeof(G->payload_dir), "%s", d);
      }
      G->attached[0] = (int) getpid();
      __sync_synchronize();
      G->magic = MM_MAGIC;
      /* wait for everybody, then release */
      for (;;)
      {
         int n = 0;
         for (i = 0; i < g_size; i++) if (G->attached[i]) n++;
         if (n == g_size) break;
         usleep(200);
      }
      G->go = 1;
   }
   else
   {
      /* attach; a stale segment of a crashed job with the same name never says "go" to us:
         re-open by name until the live one appears */
      long long t_start = now_ns();
      for (;;)
      {
         long long t_try;
         fd = shm_open(name, O_RDWR, 0600);
         if (fd < 0) { if (now_ns() - t_start > 600LL * 1000000000LL) die("shm_open(attach) timed out"); usleep(1000); continue; }
         {
            struct stat sb;
            if (fstat(fd, &sb) != 0 || (size_t) sb.st_size < g_shm_bytes) { close(fd); usleep(1000); continue; }
         }
         G = (mm_shared *) mmap(NULL, g_shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
         close(fd);
         if (G == MAP_FAILED) die("mmap(attach)");
         t_try = now_ns();
         while (G->magic != MM_MAGIC && now_ns() - t_try < 2000000000LL) usleep(200);
         if (G->magic == MM_MAGIC && G->nranks == g_size && !G->go)
         {
            G->attached[g_rank] = (int) getpid();
            while (!G->go && now_ns() - t_try < 5000000000LL) usleep(200);
            if (G->go) break;
         }
         munmap((void *) G, g_shm_bytes);
         G = NULL;
         if (now_ns() - t_start > 600LL * 1000000000LL) die("attach timed out");
         usleep(2000);
      }
   }
   dt_init();
   for (i = 0; i < g_size; i++) world[i] = i;
   ncomms = 16;
   comms = (mm_comm *) calloc(ncomms, sizeof(mm_comm));
   comms[MPI_COMM_WORLD].used = 1; comms[MPI_COMM_WORLD].ctx = 2; comms[MPI_COMM_WORLD].size = g_size;
   comms[MPI_COMM_WORLD].rank = g_rank;
   comms[MPI_COMM_WORLD].ranks = (int *) malloc(sizeof(int) * g_size);
   memcpy(comms[MPI_COMM_WORLD].ranks, world, sizeof(int) * g_size);
   comms[MPI_COMM_SELF].used = 1; comms[MPI_COMM_SELF].ctx = 4; comms[MPI_COMM_SELF].size = 1; comms[MPI_COMM_SELF].rank = 0;
   comms[MPI_COMM_SELF].ranks = (int *) malloc(sizeof(int));
   comms[MPI_COMM_SELF].ranks[0] = g_rank;
   g_init = 1;
   atexit(cleanup_atexit);
   return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
   char name[200];
   if (!g_init || g_final) return MPI_SUCCESS;
   MPI_Barrier(MPI_COMM_WORLD);
   g_final = 1;
   if (g_rank == 0)
   {
      snprintf(name, sizeof(name), "/minimpi_%s", g_job);
      shm_unlink(name);
   }
   return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode)
{
   (void) comm;
   fprintf(stderr, "[minimpi rank %d] MPI_Abort(%d)\n", g_rank, errorcode);
   fflush(stderr);
   if (G) { int i; for (i = 0; i < g_size; i++) if (i != g_rank && G->attached[i] > 0) kill(G->attached[i], SIGTERM); }
   cleanup_atexit();
   _exit(errorcode ? errorcode : 1);
}

double MPI_Wtime(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}
double MPI_Wtick(void) { return 1e-9; }
