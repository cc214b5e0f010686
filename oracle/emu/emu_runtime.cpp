// oracle/emu/emu_runtime.cpp — TEST INFRASTRUCTURE: the fiber scheduler behind cuda_runtime.h.
//
// launch(grid, block, smem, body): for every block, `block` fibers run `body` (the kernel call
// with its arguments bound).  A fiber runs until it finishes or reaches __syncthreads() / a warp
// shuffle, where it waits for the other threads of its block / warp exactly as on the device;
// threads that have left the kernel count as arrived.  One OS thread, round-robin, deterministic.
#include "cuda_runtime.h"
#include <string>
#include <fcntl.h>
#include <map>
#include <signal.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>
#include <vector>

namespace hb_emu {

ThreadCtx *g_cur = nullptr;

namespace {

constexpr size_t kStack = 64 * 1024;
constexpr int    kMaxThreads = 1024;

struct Fiber {
   ucontext_t ctx;
   ThreadCtx  tc;
   bool       done = false;
};

struct Block {
   int nthreads = 0, live = 0;
   // block barrier
   unsigned bar_gen = 0;
   int      bar_arrived = 0;
   // warp exchange (two barriers per shuffle: publish, then read)
   unsigned wgen[kMaxThreads / 32];
   int      warrived[kMaxThreads / 32];
   int      wlive[kMaxThreads / 32];
   unsigned long long wslot[kMaxThreads / 32][32];
   unsigned char      wvalid[kMaxThreads / 32][32];   // lanes taking part in the vote in flight
};

ucontext_t               g_sched;
std::vector<Fiber>       g_fibers;
std::vector<char>        g_stacks;
Block                    g_blk;
int                      g_running = -1;
const std::function<void()> *g_body = nullptr;
std::vector<char>        g_dyn;
long long                g_launches = 0;
unsigned long long       g_events = 0;      // barrier releases + thread exits: progress of a block

void yield_to_scheduler()
{
   Fiber &f = g_fibers[(size_t) g_running];
   swapcontext(&f.ctx, &g_sched);
}

void retire_thread(int t)
{
   // a thread that leaves the kernel counts as arrived at every later barrier
   Block &b = g_blk;
   b.live--;
   g_events++;
   if (b.live > 0 && b.bar_arrived == b.live) { b.bar_arrived = 0; b.bar_gen++; }
   const int w = t / 32;
   b.wlive[w]--;
   if (b.wlive[w] > 0 && b.warrived[w] == b.wlive[w]) { b.warrived[w] = 0; b.wgen[w]++; }
}

void fiber_entry()
{
   (*g_body)();
   const int t = g_running;
   g_fibers[(size_t) t].done = true;
   retire_thread(t);
   yield_to_scheduler();
}

void warp_barrier(int w)
{
   Block &b = g_blk;
   const unsigned gen = b.wgen[w];
   b.warrived[w]++;
   if (b.warrived[w] == b.wlive[w]) { b.warrived[w] = 0; b.wgen[w]++; g_events++; return; }
   while (b.wgen[w] == gen) yield_to_scheduler();
}

}  // namespace

void syncthreads()
{
   Block &b = g_blk;
   const unsigned gen = b.bar_gen;
   b.bar_arrived++;
   if (b.bar_arrived == b.live) { b.bar_arrived = 0; b.bar_gen++; g_events++; return; }
   while (b.bar_gen == gen) yield_to_scheduler();
}

// __syncwarp: every live lane of the warp arrives before any of them goes on
void syncwarp() { warp_barrier(g_running / 32); }

unsigned long long shfl_down_bits(unsigned long long bits, unsigned delta, int width)
{
   const int t = g_running, w = t / 32, lane = t % 32;
   g_blk.wslot[w][lane] = bits;
   warp_barrier(w);
   // lanes of one `width`-wide segment exchange among themselves; out of range -> own value
   const int seg_end = (lane / width) * width + width;
   const int src = lane + (int) delta;
   const unsigned long long out = (src < seg_end && src < 32) ? g_blk.wslot[w][src] : bits;
   warp_barrier(w);
   return out;
}

// warp vote: every lane of the warp publishes its predicate; all = AND over the lanes that are still in the kernel
int vote_all(int pred)
{
   const int t = g_running, w = t / 32, lane = t % 32;
   g_blk.wslot[w][lane] = pred ? 1ull : 0ull;
   g_blk.wvalid[w][lane] = 1;
   warp_barrier(w);
   int all = 1;
   for (int l = 0; l < 32; l++) { if (g_blk.wvalid[w][l] && !g_blk.wslot[w][l]) all = 0; }
   warp_barrier(w);
   g_blk.wvalid[w][lane] = 0;
   return all;
}

void *dyn_smem() { return g_dyn.data(); }
long long launches() { return g_launches; }

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body)
{
   if (grid == 0 || block == 0) return;
   if (block > (unsigned) kMaxThreads || (block % 32) != 0) {
      fprintf(stderr, "hb_emu: unsupported block size %u\n", block);
      abort();
   }
   if (g_running >= 0) {
      fprintf(stderr, "hb_emu: nested kernel launch\n");
      abort();
   }
   g_launches++;
   if (g_fibers.size() < block) g_fibers.resize(block);
   if (g_stacks.size() < (size_t) block * kStack) g_stacks.resize((size_t) block * kStack);
   if (g_dyn.size() < smem + 64) g_dyn.resize(smem + 64);
   g_body = &body;
   for (unsigned bid = 0; bid < grid; bid++) {
      Block &b = g_blk;
      b.nthreads = b.live = (int) block;
      b.bar_gen = 0; b.bar_arrived = 0;
      for (unsigned w = 0; w < block / 32; w++) { b.wgen[w] = 0; b.warrived[w] = 0; b.wlive[w] = 32; memset(b.wvalid[w], 0, 32); }
      for (unsigned t = 0; t < block; t++) {
         Fiber &f = g_fibers[t];
         f.done = false;
         f.tc.tid.x = t; f.tc.bid.x = bid; f.tc.bdim.x = block; f.tc.gdim.x = grid;
         getcontext(&f.ctx);
         f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t) t * kStack;
         f.ctx.uc_stack.ss_size = kStack;
         f.ctx.uc_link = &g_sched;
         makecontext(&f.ctx, (void (*)()) fiber_entry, 0);
      }
      int remaining = (int) block;
      while (remaining > 0) {
         const unsigned long long before = g_events;
         for (unsigned t = 0; t < block; t++) {
            Fiber &f = g_fibers[t];
            if (f.done) continue;
            g_running = (int) t;
            g_cur = &f.tc;
            swapcontext(&g_sched, &f.ctx);
            if (f.done) remaining--;
         }
         if (remaining > 0 && g_events == before) {
            fprintf(stderr, "hb_emu: block %u deadlocked (%d threads waiting at a barrier nobody else reaches)\n", bid,
                    remaining);
            abort();
         }
      }
      g_running = -1;
      g_cur = nullptr;
   }
   g_body = nullptr;
}

// ---------------------------------------------------------------------------------------
// stream capture and graphs: a graph is the list of recorded operations
// ---------------------------------------------------------------------------------------
namespace {
struct Graph { std::vector<std::function<void()>> ops; };
Graph *g_capture = nullptr;
}  // namespace

bool capturing() { return g_capture != nullptr; }
void record(std::function<void()> op) { g_capture->ops.push_back(std::move(op)); }

int capture_begin()
{
   if (g_capture) return 1;
   g_capture = new Graph();
   return 0;
}

int capture_end(void **graph)
{
   if (!g_capture) { *graph = nullptr; return 1; }
   *graph = g_capture;
   g_capture = nullptr;
   return 0;
}

int graph_instantiate(void **exec, void *graph)
{
   if (!graph) return 1;
   *exec = new Graph(*(Graph *) graph);
   return 0;
}

int graph_launch(void *exec)
{
   if (!exec || g_capture) return 1;
   for (auto &op : ((Graph *) exec)->ops) op();
   return 0;
}

void graph_destroy(void *g) { delete (Graph *) g; }

// ---------------------------------------------------------------------------------------
// "device" memory and its export to other processes
// ---------------------------------------------------------------------------------------
namespace {
std::map<void *, size_t> g_big;        // page-mapped allocations (>= 64 MB) and opened imports
std::vector<std::string> g_shm_names;
int g_shm_seq = 0;
struct IpcHandle { char name[48]; unsigned long long size; };
size_t page_round(size_t n) { return (n + 4095) & ~(size_t) 4095; }
void unlink_all() { for (auto &n : g_shm_names) shm_unlink(n.c_str()); }
// a job stopped by `timeout` / a test harness (SIGTERM, SIGINT, SIGHUP) must not leave its arena (256 MB of tmpfs) behind
void unlink_on_signal(int sig)
{
   unlink_all();
   signal(sig, SIG_DFL);
   raise(sig);
}
}  // namespace

void *dev_alloc(size_t bytes)
{
   if (bytes < (64u << 20)) return malloc(bytes);     // only the exportable arena-sized blocks are page-mapped
   const size_t n = page_round(bytes);
   void *p = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
   if (p == MAP_FAILED) return nullptr;
   g_big[p] = n;
   return p;
}

void dev_free(void *p)
{
   if (!p) return;
   auto it = g_big.find(p);
   if (it == g_big.end()) { free(p); return; }
   munmap(p, it->second);
   g_big.erase(it);
}

int ipc_export(void *handle64, void *p)
{
   auto it = g_big.find(p);
   if (it == g_big.end()) return 1;                     // only page-mapped blocks can be exported
   IpcHandle h;
   memset(&h, 0, sizeof(h));
   snprintf(h.name, sizeof(h.name), "/hb_emu_%d_%d", (int) getpid(), g_shm_seq++);
   h.size = it->second;
   const int fd = shm_open(h.name, O_CREAT | O_EXCL | O_RDWR, 0600);
   if (fd < 0) return 1;
   if (g_shm_names.empty()) {
      atexit(unlink_all);
      signal(SIGTERM, unlink_on_signal); signal(SIGINT, unlink_on_signal); signal(SIGHUP, unlink_on_signal);
   }
   g_shm_names.push_back(h.name);
   if (ftruncate(fd, (off_t) h.size) != 0) { close(fd); return 1; }
   // keep the contents and the address: copy out, map the shared object over the block, copy back
   std::vector<char> keep((char *) p, (char *) p + h.size);
   void *q = mmap(p, h.size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
   close(fd);
   if (q != p) return 1;
   memcpy(p, keep.data(), h.size);
   static_assert(sizeof(IpcHandle) <= 64, "handle fits cudaIpcMemHandle_t");
   memcpy(handle64, &h, sizeof(h));
   return 0;
}

int ipc_open(void **p, const void *handle64)
{
   IpcHandle h;
   memcpy(&h, handle64, sizeof(h));
   const int fd = shm_open(h.name, O_RDWR, 0600);
   if (fd < 0) return 1;
   void *q = mmap(nullptr, h.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
   close(fd);
   if (q == MAP_FAILED) return 1;
   g_big[q] = h.size;
   *p = q;
   return 0;
}

int ipc_close(void *p) { dev_free(p); return 0; }

}  // namespace hb_emu
