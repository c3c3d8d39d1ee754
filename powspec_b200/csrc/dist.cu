// One mesh over several GPUs: x-slab decomposition driven from inside the library
// (the reference has one address space, src/genr_mesh.c:650-747, so
// this is new design).  Rank r owns the x-planes [r Ng/G, (r+1) Ng/G) of every real
// field.  Per catalogue:
//
//   route      particles -> owner of their base x-cell         all-to-all-v
//   assign     scatter into the slab buffer (owned + halos)    assign.cu kernels
//   halo       1 plane down, 3 planes up, added by the owner   ring send/recv
//   fft (z,y)  r2c rows + strided y pass whose store writes the transpose's send
//              layout (fft_strided.cu, FftOut) — or, with peer memory, straight into
//              the destination rank's buffer: FFT and transpose in ONE kernel
//   transpose  (x-slab, y, k) -> (x, y-slab, k)                 all-to-all  | none (peer stores)
//   fft (x)    strided pass on the y-slab
//   bin        fused combine / window / L_l(mu) / reduce        binning.cu
//   reduce     nl * nbin power sums                             allreduce
//
// Everything is enqueued on the context's stream (transposes on a second stream, so
// field 1's z/y passes overlap field 0's transpose and field 0's x pass overlaps field
// 1's); the host blocks once per particle chunk (routing counts) and once at the end.
//
// Two transports carry the exchanges:
//   * NCCL (one process per GPU: bench.py under torchrun; libnccl is dlopen'ed, the
//     library has no link-time dependency on it), and
//   * "local": the ranks are threads of ONE process (what the reference's C host gets
//     through genr_mesh()/powspec(), POWSPEC_B200_DEVICES=0,1,..; also several virtual
//     ranks on one GPU for the tests): peer copies + cross-stream events, no NCCL.

#include "psb_context.h"

#if __has_include(<nccl.h>)
#include <nccl.h>
#else   // the handful of declarations used below (stable since NCCL 2.7)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4,
  ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
#endif
#include <dlfcn.h>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace psb;
using namespace psb_host;

namespace {

// ---------------------------------------------------------------------------
// NCCL, resolved at run time
// ---------------------------------------------------------------------------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
      cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy a host framework (torch) has already mapped, else POWSPEC_B200_NCCL, else
    // the loader's search path
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) if (const char *p = getenv("POWSPEC_B200_NCCL")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.lib = h;
#define PSB_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name))
    PSB_SYM(GetUniqueId); PSB_SYM(CommInitRank); PSB_SYM(CommDestroy); PSB_SYM(GroupStart);
    PSB_SYM(GroupEnd); PSB_SYM(Send); PSB_SYM(Recv); PSB_SYM(AllReduce); PSB_SYM(AllGather);
    PSB_SYM(GetErrorString);
#undef PSB_SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GroupStart || !api.GroupEnd ||
        !api.Send || !api.Recv || !api.AllReduce || !api.AllGather)
      api.lib = nullptr;
  });
  return api.lib ? &api : nullptr;
}

#define PSB_NCCL(call)                                                          \
  do {                                                                          \
    ncclResult_t r_ = (call);                                                   \
    if (r_ != ncclSuccess) {                                                    \
      set_error("NCCL failure %s at %s:%d: %s\n", #call, __FILE__, __LINE__,    \
          api->GetErrorString ? api->GetErrorString(r_) : "?");                 \
      return -1;                                                                \
    }                                                                           \
  } while (0)

// ---------------------------------------------------------------------------
// transports: stream-ordered collectives over bytes
// ---------------------------------------------------------------------------
struct Transport {
  int nranks = 1, rank = 0;
  virtual ~Transport() {}
  virtual const char *name() const = 0;
  // block q of `send` goes to rank q, block q of `recv` comes from rank q
  virtual int alltoall(const void *send, void *recv, size_t block, cudaStream_t st) = 0;
  virtual int alltoallv(const void *send, const size_t *sbytes, const size_t *sdisp, void *recv,
      const size_t *rbytes, const size_t *rdisp, cudaStream_t st) = 0;
  // ring: to_prev -> rank-1 (arrives there as from_next), to_next -> rank+1 (from_prev)
  virtual int halo(const void *to_prev, void *from_next, size_t lo_bytes, const void *to_next,
      void *from_prev, size_t hi_bytes, cudaStream_t st) = 0;
  virtual int allreduce_sum(double *buf, size_t n, cudaStream_t st) = 0;
  virtual int allgather(const void *mine, void *all, size_t bytes, cudaStream_t st) = 0;
  // every rank's buffer `mine` as a pointer this rank's kernels can store through
  // (peer memory); -1 if the ranks cannot map each other's memory
  virtual int peer_pointers(void *mine, void **all) { (void) mine; (void) all; return -1; }
  // all ranks' work enqueued before this point is complete before anything enqueued after it
  virtual int barrier(cudaStream_t st) = 0;
};

__global__ void k_sum_blocks(const double *__restrict__ parts, int nparts, size_t n, double *__restrict__ out) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    double s = 0;
    for (int q = 0; q < nparts; q++) s += parts[(size_t) q * n + i];   // rank order: same bits everywhere
    out[i] = s;
  }
}

// ---- ranks = threads of one process ---------------------------------------
struct LocalHub {
  int n;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long gen = 0;
  bool failed = false;
  struct Slot {
    int device = 0;
    const void *a = nullptr, *b = nullptr;
    const size_t *disp = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
  };
  std::vector<Slot> slot;
  explicit LocalHub(int n_) : n(n_), slot(n_) {}
  // returns false if some rank has failed (nobody waits for a rank that gave up)
  bool barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (failed) return false;
    const unsigned long g = gen;
    if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g || failed; });
    return !failed;
  }
  void fail() {
    std::lock_guard<std::mutex> lk(m);
    failed = true;
    cv.notify_all();
  }
};

struct LocalTransport : Transport {
  std::shared_ptr<LocalHub> hub;
  int device;
  DevBuf tmp;
  LocalTransport(std::shared_ptr<LocalHub> h, int r, int dev) : hub(std::move(h)), device(dev) {
    nranks = hub->n; rank = r;
    LocalHub::Slot &s = hub->slot[r];
    s.device = dev;
    cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
  }
  ~LocalTransport() override {
    LocalHub::Slot &s = hub->slot[rank];
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.done) cudaEventDestroy(s.done);
    s.ready = s.done = nullptr;
    tmp.release();
  }
  const char *name() const override { return "local"; }
  int copy_from(int q, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    if (!bytes) return 0;
    if (hub->slot[q].device == device) PSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    else PSB_CUDA(cudaMemcpyPeerAsync(dst, device, src, hub->slot[q].device, bytes, st));
    return 0;
  }
  // publish (a, b, disp) and mark this rank's inputs ready on st
  int open(const void *a, const void *b, const size_t *disp, cudaStream_t st) {
    LocalHub::Slot &s = hub->slot[rank];
    s.a = a; s.b = b; s.disp = disp;
    PSB_CUDA(cudaEventRecord(s.ready, st));
    if (!hub->barrier()) { set_error("another rank failed\n"); return -1; }
    return 0;
  }
  int wait_ready(int q, cudaStream_t st) {
    if (q != rank) PSB_CUDA(cudaStreamWaitEvent(st, hub->slot[q].ready, 0));
    return 0;
  }
  // this rank has enqueued all its reads of the peers' buffers; afterwards nobody's
  // stream runs ahead until every rank's reads are done
  int close(cudaStream_t st) {
    PSB_CUDA(cudaEventRecord(hub->slot[rank].done, st));
    if (!hub->barrier()) { set_error("another rank failed\n"); return -1; }
    for (int q = 0; q < nranks; q++)
      if (q != rank) PSB_CUDA(cudaStreamWaitEvent(st, hub->slot[q].done, 0));
    // the events may be re-recorded by the next collective only after every rank has
    // enqueued the waits above
    if (!hub->barrier()) { set_error("another rank failed\n"); return -1; }
    return 0;
  }
  int alltoall(const void *send, void *recv, size_t block, cudaStream_t st) override {
    if (open(send, nullptr, nullptr, st)) return -1;
    for (int i = 0; i < nranks; i++) {
      const int q = (rank + i) % nranks;        // stagger the pulls over the peers
      if (wait_ready(q, st)) return -1;
      if (copy_from(q, static_cast<char *>(recv) + (size_t) q * block,
            static_cast<const char *>(hub->slot[q].a) + (size_t) rank * block, block, st))
        return -1;
    }
    return close(st);
  }
  int alltoallv(const void *send, const size_t *, const size_t *sdisp, void *recv, const size_t *rbytes,
      const size_t *rdisp, cudaStream_t st) override {
    if (open(send, nullptr, sdisp, st)) return -1;
    for (int i = 0; i < nranks; i++) {
      const int q = (rank + i) % nranks;
      if (wait_ready(q, st)) return -1;
      if (copy_from(q, static_cast<char *>(recv) + rdisp[q],
            static_cast<const char *>(hub->slot[q].a) + hub->slot[q].disp[rank], rbytes[q], st))
        return -1;
    }
    return close(st);
  }
  int halo(const void *to_prev, void *from_next, size_t lo_bytes, const void *to_next, void *from_prev,
      size_t hi_bytes, cudaStream_t st) override {
    if (open(to_prev, to_next, nullptr, st)) return -1;
    const int prv = (rank + nranks - 1) % nranks, nxt = (rank + 1) % nranks;
    if (wait_ready(nxt, st) || wait_ready(prv, st)) return -1;
    if (copy_from(nxt, from_next, hub->slot[nxt].a, lo_bytes, st) ||
        copy_from(prv, from_prev, hub->slot[prv].b, hi_bytes, st))
      return -1;
    return close(st);
  }
  int allreduce_sum(double *buf, size_t n, cudaStream_t st) override {
    if (tmp.reserve(sizeof(double) * n * nranks)) return -1;
    if (open(buf, nullptr, nullptr, st)) return -1;
    for (int q = 0; q < nranks; q++) {
      if (wait_ready(q, st)) return -1;
      if (copy_from(q, tmp.as<double>() + (size_t) q * n, hub->slot[q].a, sizeof(double) * n, st)) return -1;
    }
    if (close(st)) return -1;
    k_sum_blocks<<<(unsigned) std::min<size_t>((n + 255) / 256, 1024), 256, 0, st>>>(tmp.as<double>(), nranks, n, buf);
    PSB_CUDA(cudaGetLastError());
    return 0;
  }
  int allgather(const void *mine, void *all, size_t bytes, cudaStream_t st) override {
    if (open(mine, nullptr, nullptr, st)) return -1;
    for (int q = 0; q < nranks; q++) {
      if (wait_ready(q, st)) return -1;
      if (copy_from(q, static_cast<char *>(all) + (size_t) q * bytes, hub->slot[q].a, bytes, st)) return -1;
    }
    return close(st);
  }
  int peer_pointers(void *mine, void **all) override {
    hub->slot[rank].a = mine;
    if (!hub->barrier()) return -1;
    bool ok = true;
    for (int q = 0; q < nranks; q++) {
      all[q] = const_cast<void *>(hub->slot[q].a);
      if (hub->slot[q].device != device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, device, hub->slot[q].device);
        if (!can) ok = false;
        else {
          cudaError_t e = cudaDeviceEnablePeerAccess(hub->slot[q].device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
          cudaGetLastError();
        }
      }
    }
    if (!hub->barrier()) return -1;
    return ok ? 0 : -1;
  }
  int barrier(cudaStream_t st) override {
    if (open(nullptr, nullptr, nullptr, st)) return -1;
    for (int q = 0; q < nranks; q++) if (wait_ready(q, st)) return -1;
    return close(st);
  }
};

// ---- NCCL -------------------------------------------------------------------
struct NcclTransport : Transport {
  NcclApi *api;
  ncclComm_t comm = nullptr;
  DevBuf one, ipcbuf;
  std::vector<void *> opened;
  NcclTransport(NcclApi *a) : api(a) {}
  ~NcclTransport() override {
    for (void *p : opened) cudaIpcCloseMemHandle(p);
    if (comm) api->CommDestroy(comm);
    one.release(); ipcbuf.release();
  }
  const char *name() const override { return "nccl"; }
  int alltoall(const void *send, void *recv, size_t block, cudaStream_t st) override {
    PSB_NCCL(api->GroupStart());
    for (int i = 0; i < nranks; i++) {
      const int q = (rank + i) % nranks;
      PSB_NCCL(api->Send(static_cast<const char *>(send) + (size_t) q * block, block, ncclChar, q, comm, st));
      PSB_NCCL(api->Recv(static_cast<char *>(recv) + (size_t) q * block, block, ncclChar, q, comm, st));
    }
    PSB_NCCL(api->GroupEnd());
    return 0;
  }
  int alltoallv(const void *send, const size_t *sbytes, const size_t *sdisp, void *recv,
      const size_t *rbytes, const size_t *rdisp, cudaStream_t st) override {
    PSB_NCCL(api->GroupStart());
    for (int i = 0; i < nranks; i++) {
      const int q = (rank + i) % nranks;
      if (sbytes[q]) PSB_NCCL(api->Send(static_cast<const char *>(send) + sdisp[q], sbytes[q], ncclChar, q, comm, st));
      if (rbytes[q]) PSB_NCCL(api->Recv(static_cast<char *>(recv) + rdisp[q], rbytes[q], ncclChar, q, comm, st));
    }
    PSB_NCCL(api->GroupEnd());
    return 0;
  }
  int halo(const void *to_prev, void *from_next, size_t lo_bytes, const void *to_next, void *from_prev,
      size_t hi_bytes, cudaStream_t st) override {
    const int prv = (rank + nranks - 1) % nranks, nxt = (rank + 1) % nranks;
    // with two ranks both messages go to the same peer: its first receive (from_next)
    // matches this rank's first send (to_prev)
    PSB_NCCL(api->GroupStart());
    PSB_NCCL(api->Send(to_prev, lo_bytes, ncclChar, prv, comm, st));
    PSB_NCCL(api->Send(to_next, hi_bytes, ncclChar, nxt, comm, st));
    PSB_NCCL(api->Recv(from_next, lo_bytes, ncclChar, nxt, comm, st));
    PSB_NCCL(api->Recv(from_prev, hi_bytes, ncclChar, prv, comm, st));
    PSB_NCCL(api->GroupEnd());
    return 0;
  }
  int allreduce_sum(double *buf, size_t n, cudaStream_t st) override {
    PSB_NCCL(api->AllReduce(buf, buf, n, ncclFloat64, ncclSum, comm, st));
    return 0;
  }
  int allgather(const void *mine, void *all, size_t bytes, cudaStream_t st) override {
    PSB_NCCL(api->AllGather(mine, all, bytes, ncclChar, comm, st));
    return 0;
  }
  int barrier(cudaStream_t st) override {
    if (one.reserve(sizeof(double))) return -1;
    PSB_NCCL(api->AllReduce(one.p, one.p, 1, ncclFloat64, ncclSum, comm, st));
    return 0;
  }
  // CUDA IPC: `mine` must be the base of a cudaMalloc allocation
  int peer_pointers(void *mine, void **all) override {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, mine) != cudaSuccess) { cudaGetLastError(); return -1; }
    const size_t hb = sizeof h;
    if (ipcbuf.reserve(hb * (nranks + 1))) return -1;
    char *dev = ipcbuf.as<char>();
    std::vector<cudaIpcMemHandle_t> hs(nranks);
    if (cudaMemcpy(dev, &h, hb, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (allgather(dev, dev + hb, hb, nullptr)) return -1;
    if (cudaStreamSynchronize(nullptr) != cudaSuccess ||
        cudaMemcpy(hs.data(), dev + hb, hb * nranks, cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
    int ok = 1;
    for (int q = 0; q < nranks; q++) {
      if (q == rank) { all[q] = mine; continue; }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, hs[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        all[q] = nullptr;
        continue;
      }
      opened.push_back(p);
      all[q] = p;
    }
    // all or nothing: a rank that could not map a peer makes everybody fall back
    double flag = ok ? 0.0 : 1.0;
    if (one.reserve(sizeof(double))) return -1;
    cudaMemcpy(one.p, &flag, sizeof flag, cudaMemcpyHostToDevice);
    if (allreduce_sum(one.as<double>(), 1, nullptr)) return -1;
    cudaStreamSynchronize(nullptr);
    cudaMemcpy(&flag, one.p, sizeof flag, cudaMemcpyDeviceToHost);
    return flag == 0.0 ? 0 : -1;
  }
};

// hist (u32 per destination) -> byte-free counts (u64) and the scatter cursors
__global__ void k_route_counts(const uint32_t *__restrict__ hist, int nranks, uint32_t *__restrict__ cursor,
    unsigned long long *__restrict__ counts) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t acc = 0;
  for (int r = 0; r < nranks; r++) { cursor[r] = acc; counts[r] = hist[r]; acc += hist[r]; }
}

enum { D_ROUTE = 0, D_ASSIGN, D_HALO, D_FFT_ZY, D_TRANSPOSE, D_FFT_X, D_BIN, D_REDUCE, D_COUNT };

}  // namespace

// ---------------------------------------------------------------------------
// one rank
// ---------------------------------------------------------------------------
struct psb_dist {
  psb_context *c = nullptr;
  Transport *tr = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t st_comm = nullptr;
  psb_params par;
  psb_slab sl;
  AssignGeom g;
  bool begun = false;
  // The transposes: 1 = the y pass stores straight into the peers' buffers (FFT + transpose in
  // one kernel), 0 = send layout + all-to-all on the second stream, overlapped with the next
  // field's z / y passes, -1 = choose: measured on 8 B200 the SM-issued 128-byte peer stores
  // sustain ~500 GB/s and stall the FFT kernel that issues them, the all-to-all 450-560 GB/s
  // beside the compute — config 4: 262 ms with peer stores, 213 ms with the all-to-all; config
  // 2: 12.8 vs 12.5 ms — while on 2 GPUs (half the data stays local) the fused kernel wins.
  int want_p2p = -1;
  bool p2p = false;
  long runs = 0;

  DevBuf slab[2][2];            // [cat][field]: (nx + halos) planes of reals; later k-space
  DevBuf xbuf;                  // one more slab-sized buffer (send buffer / first receive buffer)
  DevBuf xbuf2;                 // second send buffer (NCCL / local transposes of two fields in flight)
  DevBuf halo_rx, sorted, recvp, counts, plsum;
  void *peer_base[5][FftOut::MAXB];     // peer_base[b][q]: buffer b of rank q (b: slab[0][0..1], slab[1][0..1], xbuf)
  bool have_peers = false;
  void *peers_of[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // the buffers the table was built for
  unsigned long long *counts_host = nullptr;    // pinned, nranks * nranks
  size_t nrouted[2] = {0, 0};

  std::vector<Interval> intervals;
  double ms[D_COUNT];
  double a2a_bytes = 0;         // bytes this rank sent through the transposes of the last run
  double route_bytes = 0;       // particle bytes this rank sent to other ranks
};

namespace {

struct DScope {
  psb_dist *d; int stage; cudaStream_t s; cudaEvent_t a;
  DScope(psb_dist *d_, int stage_, cudaStream_t s_) : d(d_), stage(stage_), s(s_) {
    a = get_event(d->c);
    cudaEventRecord(a, s);
  }
  ~DScope() {
    cudaEvent_t b = get_event(d->c);
    cudaEventRecord(b, s);
    d->intervals.push_back({stage, a, b});
  }
};

size_t slab_bytes(const psb_dist *d) {
  return psb_slab_mesh_elems(&d->par, &d->sl) * (size_t) d->par.precision;
}

void *owned_ptr(const psb_dist *d, void *buf) {
  const size_t plane = (size_t) d->g.ng * d->g.rowlen * d->par.precision;
  return static_cast<char *>(buf) + (d->nranks > 1 ? (size_t) PSB_HALO_LO * plane : 0);
}

int dist_fail(psb_dist *d) {
  // a rank that gives up must not leave the threads of the other ranks waiting
  if (d->tr && !strcmp(d->tr->name(), "local")) static_cast<LocalTransport *>(d->tr)->hub->fail();
  return -1;
}

}  // namespace

extern "C" {

int psb_dist_unique_id(void *id128) {
  NcclApi *api = nccl_api();
  if (!api) { set_error("libnccl.so.2 not found (set POWSPEC_B200_NCCL)\n"); return -1; }
  ncclUniqueId id;
  PSB_NCCL(api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

static psb_dist *dist_new(psb_context *c, Transport *tr) {
  psb_dist *d = new psb_dist();
  d->c = c; d->tr = tr; d->nranks = tr->nranks; d->rank = tr->rank;
  d->sl = {tr->nranks, tr->rank};
  for (double &m : d->ms) m = 0;
  if (const char *e = getenv("POWSPEC_B200_P2P")) d->want_p2p = atoi(e);
  if (cudaSetDevice(c->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&d->st_comm, cudaStreamNonBlocking) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void **>(&d->counts_host),
        sizeof(unsigned long long) * tr->nranks * tr->nranks, cudaHostAllocDefault) != cudaSuccess) {
    set_error("failed to set up the slab decomposition on device %d\n", c->device);
    delete tr;
    delete d;
    return nullptr;
  }
  return d;
}

// one process per GPU: every rank passes the id rank 0 obtained from psb_dist_unique_id
psb_dist *psb_dist_create_nccl(psb_context *c, int nranks, int rank, const void *id128) {
  if (!c) { set_error("no device context\n"); return nullptr; }
  NcclApi *api = nccl_api();
  if (!api) { set_error("libnccl.so.2 not found (set POWSPEC_B200_NCCL)\n"); return nullptr; }
  if (nranks < 1 || rank < 0 || rank >= nranks || nranks > FftOut::MAXB) {
    set_error("invalid slab decomposition: rank %d of %d\n", rank, nranks);
    return nullptr;
  }
  if (cudaSetDevice(c->device) != cudaSuccess) { set_error("cudaSetDevice failed\n"); return nullptr; }
  NcclTransport *tr = new NcclTransport(api);
  tr->nranks = nranks; tr->rank = rank;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclResult_t r = api->CommInitRank(&tr->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s\n", api->GetErrorString ? api->GetErrorString(r) : "?");
    delete tr;
    return nullptr;
  }
  return dist_new(c, tr);
}

void psb_dist_destroy(psb_dist *d) {
  if (!d) return;
  cudaSetDevice(d->c->device);
  cudaDeviceSynchronize();
  for (auto &iv : d->intervals) { d->c->evpool.push_back(iv.a); d->c->evpool.push_back(iv.b); }
  for (int i = 0; i < 2; i++) for (int f = 0; f < 2; f++) d->slab[i][f].release();
  d->xbuf.release(); d->xbuf2.release(); d->halo_rx.release(); d->sorted.release(); d->recvp.release();
  d->counts.release(); d->plsum.release();
  if (d->counts_host) cudaFreeHost(d->counts_host);
  if (d->st_comm) cudaStreamDestroy(d->st_comm);
  delete d->tr;
  delete d;
}

int psb_dist_set_option(psb_dist *d, const char *name, long value) {
  if (!d || !name) return -1;
  if (!strcmp(name, "p2p")) { d->want_p2p = (int) value; d->have_peers = false; return 0; }
  return psb_set_option(d->c, name, value);
}

// Start a run: geometry, buffers (zeroed), the particle-independent mode counting on
// the side stream.  Collective.
int psb_dist_begin(psb_dist *d, const psb_params *par) {
  if (!d) { set_error("no slab decomposition\n"); return -1; }
  psb_context *c = d->c;
  PSB_CUDA(cudaSetDevice(c->device));
  if (slab_geom(c, par, &d->sl, d->g)) return dist_fail(d);
  d->par = *par;
  reset_timings(c);
  c->launches = 0;
  for (auto &iv : d->intervals) { c->evpool.push_back(iv.a); c->evpool.push_back(iv.b); }
  d->intervals.clear();
  for (double &m : d->ms) m = 0;
  d->a2a_bytes = d->route_bytes = 0;
  d->nrouted[0] = d->nrouted[1] = 0;
  const size_t sb = slab_bytes(d);
  const int nf = par->intlace ? 2 : 1;
  for (int i = 0; i < par->ncat; i++)
    for (int f = 0; f < nf; f++) {
      if (d->slab[i][f].reserve(sb)) return dist_fail(d);
      PSB_CUDA(cudaMemsetAsync(d->slab[i][f].p, 0, sb, c->st));
    }
  c->bounds_used = 0;
  // a failure here (no k bin below the Nyquist frequency, ...) belongs to powspec() in
  // the reference: psb_dist_finish reports it
  c->bins_ready = false;
  errors_quiet(true);
  prepare_bins(c, par);
  errors_quiet(false);
  d->begun = true;
  return 0;
}

// Route one chunk of this rank's share of catalogue `cat` (device memory, any spatial
// distribution) to the owners of the base x-cells and scatter what arrives.  Collective:
// every rank calls it the same number of times (n may be 0).
int psb_dist_add(psb_dist *d, int cat, const double *particles, size_t n) {
  if (!d || !d->begun) { set_error("psb_dist_begin has not been called\n"); return -1; }
  psb_context *c = d->c;
  const psb_params *par = &d->par;
  PSB_CUDA(cudaSetDevice(c->device));
  if (cat < 0 || cat >= par->ncat) { set_error("no such catalogue: %d\n", cat); return dist_fail(d); }
  void *m0 = d->slab[cat][0].p, *m1 = par->intlace ? d->slab[cat][1].p : nullptr;
  const int G = d->nranks;
  if (G == 1) {
    DScope sc(d, D_ASSIGN, c->st);
    if (assign_catalog(c, particles, n, d->g, par->assign, par->precision, 1.0, m0, m1, true)) return dist_fail(d);
    d->nrouted[cat] += n;
    return 0;
  }
  if (n > 0xffffffffull) { set_error("too many particles in one psb_dist_add call\n"); return dist_fail(d); }
  size_t sbytes[FftOut::MAXB], sdisp[FftOut::MAXB], rbytes[FftOut::MAXB], rdisp[FftOut::MAXB], nrecv = 0;
  {
    DScope sc(d, D_ROUTE, c->st);
    if (c->keys.reserve((n ? n : 1) * 4) || c->hist.reserve(64 * 4) || c->cursor.reserve(64 * 4) ||
        d->counts.reserve(sizeof(unsigned long long) * (size_t) G * (G + 1)) || d->sorted.reserve((n ? n : 1) * 32))
      return dist_fail(d);
    unsigned long long *mine = d->counts.as<unsigned long long>(), *all = mine + G;
    PSB_CUDA(cudaMemsetAsync(c->hist.p, 0, 64 * 4, c->st));
    if (n && launch_owner_keys(particles, n, d->g, G, c->keys.as<uint32_t>(), c->hist.as<uint32_t>(), c->st))
      return dist_fail(d);
    k_route_counts<<<1, 32, 0, c->st>>>(c->hist.as<uint32_t>(), G, c->cursor.as<uint32_t>(), mine);
    PSB_CUDA(cudaGetLastError());
    if (launch_owner_scatter(particles, n, c->keys.as<uint32_t>(), c->cursor.as<uint32_t>(), G,
          d->sorted.as<double>(), c->st))
      return dist_fail(d);
    if (d->tr->allgather(mine, all, sizeof(unsigned long long) * G, c->st)) return dist_fail(d);
    PSB_CUDA(cudaMemcpyAsync(d->counts_host, all, sizeof(unsigned long long) * G * G, cudaMemcpyDeviceToHost, c->st));
    PSB_CUDA(cudaStreamSynchronize(c->st));       // the one host wait of the chunk: buffer sizes
    size_t so = 0;
    for (int q = 0; q < G; q++) {
      sbytes[q] = (size_t) d->counts_host[(size_t) d->rank * G + q] * 32;   // what this rank sends to q
      rbytes[q] = (size_t) d->counts_host[(size_t) q * G + d->rank] * 32;   // what q sends here
      sdisp[q] = so; so += sbytes[q];
      rdisp[q] = nrecv * 32; nrecv += rbytes[q] / 32;
      if (q != d->rank) d->route_bytes += (double) sbytes[q];
    }
    if (d->recvp.reserve((nrecv ? nrecv : 1) * 32)) return dist_fail(d);
    if (d->tr->alltoallv(d->sorted.p, sbytes, sdisp, d->recvp.p, rbytes, rdisp, c->st)) return dist_fail(d);
    c->launches += 4;
  }
  {
    DScope sc(d, D_ASSIGN, c->st);
    if (assign_catalog(c, d->recvp.as<double>(), nrecv, d->g, par->assign, par->precision, 1.0, m0, m1, true))
      return dist_fail(d);
  }
  d->nrouted[cat] += nrecv;
  return 0;
}

// The same from HOST memory (pinned or pageable): this rank's share is cut into
// `nchunks` pieces (the same number on every rank — the call is collective), uploaded
// on the copy stream into two alternating device buffers while the previous piece is
// being routed and scattered.
int psb_dist_add_host(psb_dist *d, int cat, const double *particles_host, size_t n, int nchunks) {
  if (!d || !d->begun) { set_error("psb_dist_begin has not been called\n"); return -1; }
  psb_context *c = d->c;
  PSB_CUDA(cudaSetDevice(c->device));
  if (nchunks < 1) nchunks = 1;
  const size_t per = (n + nchunks - 1) / nchunks;
  const bool pinned = n && is_pinned(particles_host);
  for (int s = 0; s < 2; s++) {
    if (c->chunkbuf[s].reserve((per ? per : 1) * 32)) return dist_fail(d);
    if (!c->ev_filled[s]) {
      PSB_CUDA(cudaEventCreateWithFlags(&c->ev_filled[s], cudaEventDisableTiming));
      PSB_CUDA(cudaEventCreateWithFlags(&c->ev_consumed[s], cudaEventDisableTiming));
    }
  }
  for (int k = 0; k < nchunks; k++) {
    const size_t lo = std::min(n, per * k), hi = std::min(n, lo + per), len = hi - lo;
    const int s = k & 1;
    double *buf = c->chunkbuf[s].as<double>();
    if (len) {
      PSB_CUDA(cudaStreamWaitEvent(c->st_copy, c->ev_consumed[s], 0));
      {
        StageScope scope(c, PSB_T_H2D, c->st_copy);
        if (h2d_async(c, buf, particles_host + 4 * lo, len * 32, pinned, c->st_copy)) return dist_fail(d);
      }
      PSB_CUDA(cudaEventRecord(c->ev_filled[s], c->st_copy));
      PSB_CUDA(cudaStreamWaitEvent(c->st, c->ev_filled[s], 0));
    }
    if (psb_dist_add(d, cat, buf, len)) return -1;
    // the chunk is read by the partition kernels only, which psb_dist_add has waited for
    // (G > 1) or which precede this record on the stream (G == 1)
    if (len) PSB_CUDA(cudaEventRecord(c->ev_consumed[s], c->st));
  }
  return 0;
}

}  // extern "C"

namespace {

// peers' buffers for the fused FFT + transpose (all ranks must agree: collective)
int setup_peers(psb_dist *d) {
  void *bufs[5] = {d->slab[0][0].p, d->slab[0][1].p, d->slab[1][0].p, d->slab[1][1].p, d->xbuf.p};
  // buffers are (re)allocated in lock-step on all ranks (same parameters), so every rank
  // takes the same branch here
  if (d->have_peers && !memcmp(bufs, d->peers_of, sizeof bufs)) return 0;
  memcpy(d->peers_of, bufs, sizeof bufs);
  d->p2p = false;
  d->have_peers = true;
  const bool want = d->want_p2p < 0 ? (strcmp(d->tr->name(), "nccl") != 0 || d->nranks <= 2) : d->want_p2p != 0;
  if (!want || d->nranks == 1) return 0;
  bool ok = true;
  for (int b = 0; b < 5; b++) {
    for (int q = 0; q < FftOut::MAXB; q++) d->peer_base[b][q] = nullptr;
    // every rank has the same set of buffers (same parameters): a missing one is missing everywhere
    if (!bufs[b]) continue;
    if (d->tr->peer_pointers(bufs[b], d->peer_base[b])) ok = false;
  }
  d->p2p = ok;
  return 0;
}

}  // namespace

extern "C" {

// Halo exchange, the distributed transforms, binning and the reduction; every rank
// returns the same result.  wdata: GLOBAL sum of weights per catalogue
// (src/genr_mesh.c:904-909).  Collective.
psb_result *psb_dist_finish(psb_dist *d, const double wdata[2]) {
  if (!d || !d->begun) { set_error("psb_dist_begin has not been called\n"); return nullptr; }
  psb_context *c = d->c;
  const psb_params *par = &d->par;
  if (cudaSetDevice(c->device) != cudaSuccess) { set_error("cudaSetDevice failed\n"); return nullptr; }
  d->begun = false;
  auto fail = [&]() { dist_fail(d); return (psb_result *) nullptr; };
  auto hard = [&](cudaError_t e) {
    if (e != cudaSuccess) { set_error("CUDA failure: %s\n", cudaGetErrorString(e)); return true; }
    return false;
  };
  const int G = d->nranks, nc = par->ncat, nf = par->intlace ? 2 : 1, prec = par->precision;
  const int ng = d->g.ng, ngk = ng / 2 + 1, nx = d->g.nx;
  const size_t plane = (size_t) ng * d->g.rowlen * prec;        // bytes of one x-plane (real, padded)
  const size_t blk = (size_t) nx * nx * ngk * 2 * prec;         // one (source, destination) block of the transpose
  const size_t sb = slab_bytes(d);
  cudaStream_t st = c->st, sc = d->st_comm;

  // ---- halo planes to their owners (periodic ring), added there
  if (G > 1) {
    DScope scope(d, D_HALO, st);
    if (d->halo_rx.reserve((PSB_HALO_LO + PSB_HALO_HI) * plane)) return fail();
    char *rx_next = d->halo_rx.as<char>(), *rx_prev = rx_next + PSB_HALO_LO * plane;
    for (int i = 0; i < nc; i++)
      for (int f = 0; f < nf; f++) {
        char *m = d->slab[i][f].as<char>();
        if (d->tr->halo(m, rx_next, PSB_HALO_LO * plane, m + (size_t) (PSB_HALO_LO + nx) * plane, rx_prev,
              PSB_HALO_HI * plane, st))
          return fail();
        // from the next rank: its planes below its slab = my last owned planes
        if (launch_add(m + (size_t) nx * plane, rx_next, PSB_HALO_LO * plane / prec, prec, st) ||
            launch_add(m + (size_t) PSB_HALO_LO * plane, rx_prev, PSB_HALO_HI * plane / prec, prec, st))
          return fail();
        c->launches += 3;
      }
  }

  // ---- transforms.  Field t = (catalogue, field) in order; k-space result of field t:
  //   G == 1           in place in its slab buffer
  //   NCCL / local     y pass writes the send layout into xbuf / xbuf2 (alternating), the
  //                    all-to-all (second stream) delivers into the field's own slab buffer
  //   peer stores      y pass writes straight into the destination ranks' buffers: field 0
  //                    into xbuf, field t > 0 into the slab buffer of field t-1 (dead once
  //                    every rank has finished that field's y pass)
  if (slab_plans(c, ng, nx, prec)) return fail();
  if (G > 1 && d->xbuf.reserve(sb)) return fail();
  if (setup_peers(d)) return fail();
  const bool p2p = d->p2p && c->slab_own;
  if (G > 1 && !p2p && nc * nf > 1 && d->xbuf2.reserve(sb)) return fail();
  const int nt = nc * nf;
  void *kspace[4] = {nullptr, nullptr, nullptr, nullptr};
  double sent_frac = 1.0;
  cudaEvent_t ev_packed[4], ev_recv[4];
  for (int t = 0; t < nt; t++) { ev_packed[t] = get_event(c); ev_recv[t] = get_event(c); }
  auto release_events = [&]() { for (int t = 0; t < nt; t++) { c->evpool.push_back(ev_packed[t]); c->evpool.push_back(ev_recv[t]); } };
  for (int t = 0; t < nt; t++) {
    const int i = t / nf, f = t % nf;
    void *buf = d->slab[i][f].p, *owned = owned_ptr(d, buf);
    FftOut out;
    int bsel = -1;      // which of this rank's buffers receives field t (index into peer_base)
    if (G > 1) {
      out.ny = nx;
      out.outer_stride = (size_t) nx * ngk;
      if (p2p) {
        bsel = t == 0 ? 4 : ((t - 1) / nf) * 2 + (t - 1) % nf;
        for (int q = 0; q < G; q++) out.base[q] = static_cast<char *>(d->peer_base[bsel][q]) + (size_t) d->rank * blk;
        kspace[t] = t == 0 ? d->xbuf.p : d->slab[(t - 1) / nf][(t - 1) % nf].p;
      }
      else {
        char *sendbuf = (t & 1) ? d->xbuf2.as<char>() : d->xbuf.as<char>();
        for (int q = 0; q < G; q++) out.base[q] = sendbuf + (size_t) q * blk;
        kspace[t] = buf;
        // the send buffer is free once the transpose two fields back has been delivered
        if (t >= 2 && hard(cudaStreamWaitEvent(st, ev_recv[t - 2], 0))) return fail();
      }
    }
    else kspace[t] = buf;
    {
      DScope scope(d, D_FFT_ZY, st);
      if (c->slab_own) {
        const int zp = c->slab_zp;
        for (int x0 = 0; x0 < nx; x0 += zp) {
          char *grp = static_cast<char *>(owned) + (size_t) x0 * plane;
          if (c->slab_own_z) {
            if (launch_fft_rows(grp, grp, prec, ng, (long) zp * ng, 2 * (size_t) ngk, ngk, st)) return fail();
          }
          else if (prec == 8) {
            if (cufftExecD2Z(c->slab_yz, (cufftDoubleReal *) grp, (cufftDoubleComplex *) grp) != CUFFT_SUCCESS) {
              set_error("cuFFT failure in the z pass\n"); return fail();
            }
          }
          else if (cufftExecR2C(c->slab_yz, (cufftReal *) grp, (cufftComplex *) grp) != CUFFT_SUCCESS) {
            set_error("cuFFT failure in the z pass\n"); return fail();
          }
        }
        if (G == 1) {
          if (launch_fft_strided(owned, prec, ng, ngk, 1, nx, nullptr, nullptr, 0.0, st)) return fail();
        }
        else {
          // rows beyond the last bin edge are neither stored nor sent (the x pass on the
          // receiving rank skips their tiles by the same test)
          FftStoreSkip ss;
          const bool sskip = c->opt_fft_skip && c->opt_fft_store_skip && c->slab_own_x && c->bins_ready;
          if (sskip) {
            if (hard(cudaStreamWaitEvent(st, c->ev_geom, 0))) return fail();
            ss.k2t = c->bg.kax2[1]; ss.k2k = c->bg.kax2[2];
            ss.k2max = c->host_tables[15 * (size_t) ng + c->nbin];
          }
          if (launch_fft_strided_out(owned, prec, ng, ngk, nx, out, st, sskip ? &ss : nullptr)) return fail();
          if (sskip && t == 0) {
            // fraction of the (y, k) rows that are stored (per column; the kernel decides per
            // tile of columns, so slightly more is sent): what really crosses the links
            const double *ky2 = &c->host_tables[(size_t) (3 + 1) * ng], *kz2 = &c->host_tables[(size_t) (3 + 2) * ng];
            size_t kept = 0;
            for (int y = 0; y < ng; y++)
              for (int k = 0; k < ngk; k++) kept += (ky2[y] + kz2[k] < ss.k2max);
            sent_frac = (double) kept / ((double) ng * ngk);
          }
        }
        c->launches += 2;
      }
      else {
        // sizes without hand-written passes: cuFFT's batched 2-D plan in place, then pack
        if (prec == 8 ? cufftExecD2Z(c->slab_yz, (cufftDoubleReal *) owned, (cufftDoubleComplex *) owned) != CUFFT_SUCCESS
                      : cufftExecR2C(c->slab_yz, (cufftReal *) owned, (cufftComplex *) owned) != CUFFT_SUCCESS) {
          set_error("cuFFT failure in the 2-D transform\n"); return fail();
        }
        c->launches++;
        if (G > 1) {
          const size_t csz = 2 * (size_t) prec, width = (size_t) nx * ngk * csz, spitch = (size_t) ng * ngk * csz;
          for (int q = 0; q < G; q++)
            if (hard(cudaMemcpy2DAsync(out.base[q], width, static_cast<const char *>(owned) + (size_t) q * width,
                    spitch, width, nx, cudaMemcpyDeviceToDevice, st)))
              return fail();
        }
      }
    }
    if (G > 1) {
      d->a2a_bytes += (double) blk * (G - 1) * (p2p ? sent_frac : 1.0);
      if (p2p) {
        // the stores into the peers are complete when every rank's kernel is: a
        // stream-ordered barrier; it also releases this field's slab buffer as the
        // next field's target
        DScope scope(d, D_TRANSPOSE, st);
        if (d->tr->barrier(st)) return fail();
      }
      else {
        if (hard(cudaEventRecord(ev_packed[t], st)) || hard(cudaStreamWaitEvent(sc, ev_packed[t], 0))) return fail();
        {
          DScope scope(d, D_TRANSPOSE, sc);
          const void *sendbuf = (t & 1) ? d->xbuf2.p : d->xbuf.p;
          if (d->tr->alltoall(sendbuf, buf, blk, sc)) return fail();
        }
        if (hard(cudaEventRecord(ev_recv[t], sc))) return fail();
      }
    }
  }
  // ---- x pass on the y-slab (the result stays transposed: binning is layout-agnostic)
  const bool bins_ok = same_bins(c, par) || !prepare_bins(c, par);
  if (!bins_ok) { release_events(); return fail(); }
  for (int t = 0; t < nt; t++) {
    if (G > 1 && !p2p && hard(cudaStreamWaitEvent(st, ev_recv[t], 0))) { release_events(); return fail(); }
    DScope scope(d, D_FFT_X, st);
    if (c->slab_own_x) {
      // columns whose smallest |k|^2 lies beyond the last bin edge are never read
      const bool skip = c->opt_fft_skip != 0;
      const double k2max = c->host_tables[15 * (size_t) ng + c->nbin];
      if (skip && hard(cudaStreamWaitEvent(st, c->ev_geom, 0))) { release_events(); return fail(); }
      FftStoreSkip ss;
      const bool sskip = skip && c->opt_fft_store_skip;
      if (sskip) {
        ss.k2t = c->bg.kax2[0]; ss.k2o = c->bg.kax2[1] + d->g.x0; ss.k2k = c->bg.kax2[2];
        ss.k2max = k2max; ss.per_column = 1;
      }
      if (launch_fft_strided(kspace[t], prec, ng, ngk, 0, nx, skip ? c->bg.kax2[1] + d->g.x0 : nullptr,
            skip ? c->bg.kax2[2] : nullptr, k2max, st, sskip ? &ss : nullptr)) { release_events(); return fail(); }
    }
    else if (prec == 8 ? cufftExecZ2Z(c->slab_x, (cufftDoubleComplex *) kspace[t], (cufftDoubleComplex *) kspace[t], CUFFT_FORWARD) != CUFFT_SUCCESS
                       : cufftExecC2C(c->slab_x, (cufftComplex *) kspace[t], (cufftComplex *) kspace[t], CUFFT_FORWARD) != CUFFT_SUCCESS) {
      set_error("cuFFT failure in the x pass\n");
      release_events();
      return fail();
    }
    c->launches++;
  }
  release_events();

  // ---- binning of the local y-slab, then one allreduce of all the sums
  const int nbin = c->nbin, nl = par->npole;
  const size_t nacc = (size_t) nl * nbin;
  if (d->plsum.reserve(sizeof(double) * (3 * nacc + 1))) return fail();
  double *pls = d->plsum.as<double>();
  if (hard(cudaMemsetAsync(pls, 0, sizeof(double) * (3 * nacc + 1), st))) return fail();
  {
    DScope scope(d, D_BIN, st);
    BinGeom bg = c->bg;
    bg.j0 = d->g.x0; bg.nj = nx;               // y-slab after the transpose: same split as x
    if (hard(cudaStreamWaitEvent(st, c->ev_geom, 0))) return fail();
    double *scratch = reinterpret_cast<double *>(c->binscratch.as<char>() + c->bin_sb);
    auto F = [&](int i, int f) { return (f < nf) ? kspace[i * nf + f] : nullptr; };
    for (int i = 0; i < nc; i++) {
      if (!par->isauto[i]) continue;
      if (launch_bin(bg, prec, F(i, 0), F(i, 1), F(i, 0), F(i, 1), pls + (size_t) i * nacc, scratch, c->bin_sb, st))
        return fail();
      c->launches += 2;
    }
    if (par->iscross && nc == 2) {
      if (launch_bin(bg, prec, F(0, 0), F(0, 1), F(1, 0), F(1, 1), pls + 2 * nacc, scratch, c->bin_sb, st))
        return fail();
      c->launches += 2;
    }
  }
  // def_box's checks for simulation boxes (src/genr_mesh.c:516-531) on the particles this
  // rank scattered; a violation anywhere fails every rank (last element of the reduction)
  std::vector<double> hb(c->bounds_used / sizeof(double));
  if (!hb.empty() && (hard(cudaMemcpyAsync(hb.data(), c->bounds_part.p, c->bounds_used, cudaMemcpyDeviceToHost, st)) ||
        hard(cudaStreamSynchronize(st))))
    return fail();
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (size_t q = 0; q + 5 < hb.size(); q += 6)
    for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], hb[q + a]); hi[a] = std::max(hi[a], hb[q + 3 + a]); }
  double bad = 0;
  int bad_axis = -1;
  for (int a = 0; a < 3; a++)
    if (lo[a] <= hi[a] && (lo[a] < 0 || hi[a] >= par->bsize[a])) { bad = 1; if (bad_axis < 0) bad_axis = a; }
  if (hard(cudaMemcpyAsync(pls + 3 * nacc, &bad, sizeof bad, cudaMemcpyHostToDevice, st))) return fail();
  if (G > 1) {
    DScope scope(d, D_REDUCE, st);
    if (d->tr->allreduce_sum(pls, 3 * nacc + 1, st)) return fail();
  }
  std::vector<double> hp(3 * nacc + 1);
  if (hard(cudaMemcpyAsync(hp.data(), pls, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, st)) ||
      hard(cudaStreamSynchronize(st)))
    return fail();
  for (auto &iv : d->intervals) {
    float t = 0;
    if (cudaEventElapsedTime(&t, iv.a, iv.b) == cudaSuccess) d->ms[iv.stage] += t;
    c->evpool.push_back(iv.a); c->evpool.push_back(iv.b);
  }
  d->intervals.clear();
  cudaGetLastError();
  d->runs++;
  if (hp[3 * nacc] != 0) {
    const char ax[3] = {'x', 'y', 'z'};
    if (bad_axis >= 0 && lo[bad_axis] < 0) set_error("%c coordinate below 0: %lf\n", ax[bad_axis], lo[bad_axis]);
    else if (bad_axis >= 0) set_error("%c coordinate not smaller than BOX_SIZE: %lf\n", ax[bad_axis], hi[bad_axis]);
    else set_error("coordinates outside the box on another rank\n");
    c->bins_ready = false;
    collect_timings(c);
    return nullptr;
  }
  const double *p0 = par->isauto[0] ? hp.data() : nullptr;
  const double *p1 = (nc == 2 && par->isauto[1]) ? hp.data() + nacc : nullptr;
  const double *px = (par->iscross && nc == 2) ? hp.data() + 2 * nacc : nullptr;
  return psb_slab_finish(c, par, p0, p1, px, wdata);
}

// stage times of the last run on this rank (CUDA events), ms: route, assign, halo,
// fft_zy (+ pack / peer stores), transpose (all-to-all on the second stream, or the
// barrier after peer stores), fft_x, bin, reduce
int psb_dist_timings(const psb_dist *d, double *ms, int n) {
  if (!d || !ms) return -1;
  for (int i = 0; i < n && i < D_COUNT; i++) ms[i] = d->ms[i];
  return std::min(n, (int) D_COUNT);
}

// what the last run moved between ranks, from this rank: [0] transpose bytes sent,
// [1] particle bytes routed away, [2] 1 if the transposes went through peer stores
int psb_dist_traffic(const psb_dist *d, double *out, int n) {
  if (!d || !out) return -1;
  const double v[3] = {d->a2a_bytes, d->route_bytes, d->p2p ? 1.0 : 0.0};
  for (int i = 0; i < n && i < 3; i++) out[i] = v[i];
  return std::min(n, 3);
}

const char *psb_dist_transport(const psb_dist *d) { return d ? d->tr->name() : ""; }
psb_context *psb_dist_context(psb_dist *d) { return d ? d->c : nullptr; }

}  // extern "C"

// ---------------------------------------------------------------------------
// all ranks in ONE process: one host thread per rank (what the reference's
// single-process C host drives through genr_mesh() / powspec())
// ---------------------------------------------------------------------------
struct psb_group {
  int n = 0;
  std::vector<int> device;
  std::shared_ptr<LocalHub> hub;
  std::vector<psb_context *> ctx;
  std::vector<psb_dist *> rank;
  std::vector<DevBuf> chunk[2];         // per rank, double-buffered upload chunks
  psb_params par;
  double wdata[2] = {0, 0};
  bool mesh_ready = false;
  long chunk_particles = 1 << 23;       // 256 MiB of records per upload chunk
};

namespace {

template <typename F> int run_ranks(psb_group *g, F body) {
  std::vector<int> rc(g->n, 0);
  std::vector<std::string> err(g->n);
  std::vector<std::thread> th;
  for (int r = 0; r < g->n; r++)
    th.emplace_back([&, r] {
      cudaSetDevice(g->device[r]);
      rc[r] = body(r);
      if (rc[r]) { err[r] = get_error(); g->hub->fail(); }
    });
  for (auto &t : th) t.join();
  for (int r = 0; r < g->n; r++)
    if (rc[r]) {
      // the first real message (not the "another rank failed" echo)
      int pick = r;
      for (int q = 0; q < g->n; q++)
        if (rc[q] && err[q].find("another rank failed") == std::string::npos) { pick = q; break; }
      set_error("%s", err[pick].c_str());
      { std::lock_guard<std::mutex> lk(g->hub->m); g->hub->failed = false; g->hub->arrived = 0; }
      return -1;
    }
  return 0;
}

}  // namespace

extern "C" {

// nranks (virtual) ranks on the listed devices; a device may appear more than once
psb_group *psb_group_create(const int *devices, int nranks) {
  if (!devices || nranks < 1 || nranks > FftOut::MAXB) { set_error("invalid device list\n"); return nullptr; }
  psb_group *g = new psb_group();
  g->n = nranks;
  g->hub = std::make_shared<LocalHub>(nranks);
  g->chunk[0].resize(nranks); g->chunk[1].resize(nranks);
  for (int r = 0; r < nranks; r++) {
    g->device.push_back(devices[r]);
    psb_context *c = psb_create(devices[r]);
    if (!c) { psb_group_destroy(g); return nullptr; }
    // pageable catalogues: every rank stages its own share (hostcopy.cpp); the ranks share the
    // host's cores, so each pool takes its part of them instead of 16 threads per rank
    const unsigned hw = std::thread::hardware_concurrency();
    c->opt_h2d_threads = std::max<long>(2, std::min<long>(c->opt_h2d_threads, (long) (hw ? hw : 16u) / nranks));
    g->ctx.push_back(c);
    psb_dist *d = dist_new(c, new LocalTransport(g->hub, r, devices[r]));
    if (!d) { psb_group_destroy(g); return nullptr; }
    g->rank.push_back(d);
  }
  return g;
}

void psb_group_destroy(psb_group *g) {
  if (!g) return;
  for (size_t r = 0; r < g->rank.size(); r++) psb_dist_destroy(g->rank[r]);
  for (size_t r = 0; r < g->ctx.size(); r++) {
    cudaSetDevice(g->device[r]);
    for (int s = 0; s < 2; s++) if (r < g->chunk[s].size()) g->chunk[s][r].release();
    psb_destroy(g->ctx[r]);
  }
  delete g;
}

int psb_group_size(const psb_group *g) { return g ? g->n : 0; }
psb_dist *psb_group_rank(psb_group *g, int r) { return (g && r >= 0 && r < g->n) ? g->rank[r] : nullptr; }

int psb_group_set_option(psb_group *g, const char *name, long value) {
  if (!g) return -1;
  if (!strcmp(name, "group_chunk")) { g->chunk_particles = std::max<long>(value, 1024); return 0; }
  int rc = 0;
  for (psb_dist *d : g->rank) rc |= psb_dist_set_option(d, name, value);
  return rc;
}

// genr_mesh() for the group: rank r takes the r-th contiguous share of every
// catalogue (host memory, or device memory of any one GPU), uploads it in chunks and
// routes / scatters it.  Simulation boxes only.
int psb_group_mesh(psb_group *g, const psb_params *par, const psb_cats *cats) {
  if (!g) { set_error("no device group\n"); return -1; }
  if (!par || !cats) { set_error("catalogs not read\n"); return -1; }
  if (cats->cnvt) { set_error("coordinate conversion is not available on the slab-decomposed path\n"); return -1; }
  g->mesh_ready = false;
  g->par = *par;
  for (int i = 0; i < 2; i++) g->wdata[i] = i < par->ncat ? cats->wdata[i] : 0;
  const size_t CH = (size_t) g->chunk_particles;
  int rc = run_ranks(g, [&](int r) -> int {
    psb_dist *d = g->rank[r];
    psb_context *c = d->c;
    if (psb_dist_begin(d, par)) return -1;
    for (int i = 0; i < par->ncat; i++) {
      const size_t N = cats->ndata[i];
      if (N && !cats->data[i]) { set_error("catalogs not read\n"); return -1; }
      const size_t per = (N + g->n - 1) / g->n;
      const size_t a = std::min(N, per * r), b = std::min(N, per * (r + 1));
      const size_t nchunk = std::max<size_t>(1, (per + CH - 1) / CH);     // the same on every rank
      if (cats->memspace == PSB_MEM_HOST) {
        if (psb_dist_add_host(d, i, cats->data[i] + 4 * a, b - a, (int) nchunk)) return -1;
        continue;
      }
      // device memory of one GPU: every rank pulls its share chunk by chunk
      for (size_t k = 0; k < nchunk; k++) {
        const size_t lo = std::min(b, a + k * CH), hi = std::min(b, lo + CH), len = hi - lo;
        DevBuf &buf = g->chunk[k & 1][r];
        if (buf.reserve(std::min(per, CH) * 32 + 32)) return -1;
        // stream order protects the buffer: its previous readers precede this copy on c->st
        if (len) PSB_CUDA(cudaMemcpyAsync(buf.p, cats->data[i] + 4 * lo, len * 32, cudaMemcpyDefault, c->st));
        if (psb_dist_add(d, i, buf.as<double>(), len)) return -1;
      }
    }
    return 0;
  });
  if (rc) return -1;
  g->mesh_ready = true;
  return 0;
}

// powspec() for the group
psb_result *psb_group_power(psb_group *g, const psb_params *par) {
  if (!g) { set_error("no device group\n"); return nullptr; }
  if (!g->mesh_ready) { set_error("meshes not generated\n"); return nullptr; }
  g->mesh_ready = false;
  (void) par;           // the binning parameters were fixed by psb_group_mesh (same CONF)
  std::vector<psb_result *> res(g->n, nullptr);
  int rc = run_ranks(g, [&](int r) -> int {
    res[r] = psb_dist_finish(g->rank[r], g->wdata);
    return res[r] ? 0 : -1;
  });
  for (int r = 1; r < g->n; r++) psb_result_free(res[r]);
  if (rc) { psb_result_free(res[0]); return nullptr; }
  return res[0];
}

psb_result *psb_group_run(psb_group *g, const psb_params *par, const psb_cats *cats) {
  if (psb_group_mesh(g, par, cats)) return nullptr;
  return psb_group_power(g, par);
}

}  // extern "C"
