// comm.cu — the one exchange step of the path: every volume lives on its own GPU (static map on rank 0, one
// ITMScene per car on the others; DS/InstRecLib/InstanceReconstructor.cpp:363-389) and per frame each rank hands the colour
// and depth renders of its volume to rank 0, which z-composites them over its own render — what the reference does with
// per-volume GetImage / GetFloatImage calls and a CPU composite (InstanceReconstructor.cpp:851-987).
//
// C++ host code. NCCL (loaded at run time: dlopen "libnccl.so.2" — a process that already carries torch's NCCL gets that one;
// libb200fusion has no link-time dependency on it and single-GPU users never touch it) sets the communicator up and carries the
// bootstrap. The per-frame hand-over itself does NOT run NCCL kernels by default:
//
//   * "push" transport (default): rank 0 exports its layer buffers and every rank a small flag block through CUDA IPC; per frame
//     a rank copies its two renders straight into rank 0's buffers with the COPY ENGINES over NVLink (cudaMemcpyAsync on
//     peer-mapped memory) and then raises a sequence flag in rank 0's memory; rank 0's stream waits on the flags with
//     cuStreamWaitValue32 (no SM), composites, and lowers the senders' "slot free" flags the same way. No rendez-vous kernel
//     ever sits on an SM and nothing competes with the frame's persistent kernels (IntegrateIntoScene and the list kernel fill
//     the register file). Measured on 2 and 4 B200s (profiles/r02_multi.md): the compositor rank's step is 179.7 us with the
//     exchange and 178.6 - 178.9 us without.
//   * "nccl" transport (B200_COMM_IMPL=nccl, and the fallback where stream memory operations are unavailable): grouped
//     ncclSend / ncclRecv (NCCL has no gather) on the communicator's stream.
//
// Everything runs on the communicator's OWN stream, ordered against the engine's stream by events only:
//   engine stream:  ... frame k's renders into slot k&1 ........ frame k+1 ........ frame k+2 (waits: slot k&1 drained)
//   comm stream:                 wait(renders k) -> send | recv xN -> composite k -> done[k&1]
// so frame k+1's kernels never wait for frame k's rendez-vous (round 1 enqueued a torch.distributed.gather on the engine
// stream every frame: a 30-60 us cross-rank barrier on the critical path, scaling efficiency 0.81-0.89).
#include "engine.h"

#include <cuda.h>
#include <dlfcn.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
typedef struct { char internal[B200_COMM_ID_BYTES]; } nccl_id_t;
typedef void *nccl_comm_t;
struct Nccl {
  void *lib;
  int (*GetUniqueId)(nccl_id_t *);
  int (*CommInitRank)(nccl_comm_t *, int, nccl_id_t, int);
  int (*CommDestroy)(nccl_comm_t);
  int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
};
Nccl g_nccl;
bool load_nccl(char *err, size_t cap) {
  if (g_nccl.lib) return true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { snprintf(err, cap, "cannot load libnccl.so.2: %s", dlerror()); return false; }
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { snprintf(err, cap, "libnccl lacks %s", name); return false; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy") SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.lib = h;
  return true;
}
thread_local char g_commErr[512] = "";

// stream memory operations of the driver API, fetched through the runtime (no link-time dependency on libcuda)
typedef CUresult (*wait_value32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
wait_value32_t g_waitValue32 = nullptr;
bool load_stream_memops() {
  if (g_waitValue32) return true;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    return false;
  }
  g_waitValue32 = (wait_value32_t)fn;
  return true;
}

// raises flags in (possibly peer-mapped) memory behind everything enqueued before it on the stream: a one-thread launch — a
// store through the peer mapping is the portable way to write another GPU's memory from a stream
struct FlagList { unsigned *p[64]; int n; };
__global__ void k_raise_flags(FlagList f, unsigned value) {
  if (threadIdx.x < f.n) { __threadfence_system(); *(volatile unsigned *)f.p[threadIdx.x] = value; __threadfence_system(); }
}
}  // namespace

struct b200_comm {
  int device, nranks, rank;
  size_t pixels;
  nccl_comm_t comm;
  cudaStream_t stream;
  cudaEvent_t evReady[2], evDone[2];
  bool busy[2];
  // rank 0: the other ranks' layers, [slot][rank-1]
  b200_vec4u *d_layerColor[2];
  float *d_layerDepth[2];
  // push transport
  bool push;
  int diag;                        // B200_COMM_DIAG (measurement / test aid): 1 = skip the composite launch, 4 = release does not wait, 8 = simulate a failed IPC open
  unsigned seq[2];                 // hand-overs submitted per slot
  char *d_block;                   // rank 0: layers + ready flags in one exported allocation; other ranks: their two "slot free" flags
  size_t blockBytes;
  unsigned *d_ready;               // rank 0: [rank-1][slot], raised by the senders (inside d_block)
  unsigned *d_free;                // ranks > 0: [slot], raised by rank 0 (== d_block)
  char *peerBlock[64];             // rank 0: rank r's flag block; rank r: rank 0's block (index 0)
  // The hand-over is a dozen stream operations per frame. They are issued by the communicator's OWN host thread, so the thread
  // that drives the engine pays one event record and a queue push per frame.
  struct Work {
    int slot; unsigned seq;
    const b200_vec4u *d_color; const float *d_depth; b200_vec4u *d_out_color; float *d_out_depth;
    int32_t tints[64 * 4]; bool hasTints; float dim_factor, tint_strength;
    b200_engine *e;
  };
  std::thread *worker;
  std::mutex *mu;
  std::condition_variable *cv;
  std::deque<Work> *queue;
  bool stop;
  unsigned enqueued[2];            // seq of the last hand-over of the slot whose stream operations have all been issued
  b200_status workerStatus;        // first failure of the worker (sticky)
  char err[512];
};

#define CCK(call)                                                                                                  \
  do {                                                                                                             \
    cudaError_t _e = (call);                                                                                       \
    if (_e != cudaSuccess) { snprintf(c->err, sizeof(c->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); return B200_ERR_CUDA; } \
  } while (0)
#define NCK(call)                                                                                                  \
  do {                                                                                                             \
    int _r = (call);                                                                                               \
    if (_r != 0) { snprintf(c->err, sizeof(c->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r)); return B200_ERR_CUDA; } \
  } while (0)

extern "C" {
static void comm_worker(b200_comm *c);

b200_status b200_comm_unique_id(char id[B200_COMM_ID_BYTES]) {
  if (!id) return B200_ERR_INVALID;
  if (!load_nccl(g_commErr, sizeof(g_commErr))) return B200_ERR_UNSUPPORTED;
  nccl_id_t u;
  const int r = g_nccl.GetUniqueId(&u);
  if (r != 0) { snprintf(g_commErr, sizeof(g_commErr), "ncclGetUniqueId -> %s", g_nccl.GetErrorString(r)); return B200_ERR_CUDA; }
  memcpy(id, u.internal, B200_COMM_ID_BYTES);
  return B200_OK;
}

void b200_comm_destroy(b200_comm *c) {
  if (!c) return;
  if (c->worker) {
    { std::lock_guard<std::mutex> lk(*c->mu); c->stop = true; }
    c->cv->notify_all();
    c->worker->join();
    delete c->worker;
  }
  delete c->queue; delete c->cv; delete c->mu;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  for (int s = 0; s < 2; ++s) {
    if (c->evReady[s]) cudaEventDestroy(c->evReady[s]);
    if (c->evDone[s]) cudaEventDestroy(c->evDone[s]);
    if (!c->push) { cudaFree(c->d_layerColor[s]); cudaFree(c->d_layerDepth[s]); }
  }
  for (int r = 0; r < 64; ++r) if (c->peerBlock[r]) cudaIpcCloseMemHandle(c->peerBlock[r]);
  cudaFree(c->d_block);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaGetLastError();
  delete c;
}

static b200_status comm_create_body(b200_comm *c, const char *id) {
  if (!load_nccl(c->err, sizeof(c->err))) return B200_ERR_UNSUPPORTED;
  CCK(cudaSetDevice(c->device));
  int lo = 0, hi = 0;
  CCK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CCK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));    // its few small launches go ahead of the frame's kernels
  const char *impl = getenv("B200_COMM_IMPL");
  { const char *dg = getenv("B200_COMM_DIAG"); c->diag = dg ? atoi(dg) : 0; }
  int memops = 0;
  cudaDeviceGetAttribute(&memops, cudaDevAttrMemSyncDomainCount, c->device);   // (any query: make sure the context exists before the driver entry point is fetched)
  c->push = c->nranks > 1 && c->nranks <= 64 && !(impl && impl[0] == 'n') && load_stream_memops();
  for (int s = 0; s < 2; ++s) {
    CCK(cudaEventCreateWithFlags(&c->evReady[s], cudaEventDisableTiming));
    CCK(cudaEventCreateWithFlags(&c->evDone[s], cudaEventDisableTiming));
    if (!c->push && c->rank == 0 && c->nranks > 1) {
      CCK(cudaMalloc(&c->d_layerColor[s], sizeof(b200_vec4u) * c->pixels * (size_t)(c->nranks - 1)));
      CCK(cudaMalloc(&c->d_layerDepth[s], sizeof(float) * c->pixels * (size_t)(c->nranks - 1)));
    }
  }
  nccl_id_t u;
  memcpy(u.internal, id, B200_COMM_ID_BYTES);
  NCK(g_nccl.CommInitRank(&c->comm, c->nranks, u, c->rank));
  if (!c->push) return B200_OK;
  // ---- push transport: one exported allocation per rank, handles exchanged over the communicator ----
  // Every CUDA step that can fail on a machine without peer access between the GPUs (export, open) only lowers `ok`; the ranks
  // then agree on the verdict over the communicator and, if any of them failed, ALL fall back to the nccl transport.
  const size_t n = c->pixels, layers = (size_t)(c->nranks - 1);
  const size_t colBytes = sizeof(b200_vec4u) * n * layers, depBytes = sizeof(float) * n * layers;
  const size_t flagBytes = 256;      // rank 0: ready[rank-1][slot] (at most 63 x 2 words); others: free[slot]
  c->blockBytes = c->rank == 0 ? 2 * (colBytes + depBytes) + 4 * flagBytes : flagBytes;
  int ok = 1;
  struct BootRec { cudaIpcMemHandle_t handle; int ok; int pad[3]; };
  BootRec mine;
  memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&c->d_block, c->blockBytes) != cudaSuccess || cudaMemset(c->d_block, 0, c->blockBytes) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine.handle, c->d_block) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  mine.ok = ok;
  const size_t hb = sizeof(BootRec);
  char *d_boot = nullptr;
  CCK(cudaMalloc(&d_boot, hb * (size_t)c->nranks));
  CCK(cudaMemcpy(d_boot + hb * (size_t)c->rank, &mine, hb, cudaMemcpyHostToDevice));
  NCK(g_nccl.GroupStart());
  if (c->rank == 0) {
    for (int r = 1; r < c->nranks; ++r) {
      NCK(g_nccl.Send(d_boot, hb, 0 /* ncclInt8 */, r, c->comm, c->stream));
      NCK(g_nccl.Recv(d_boot + hb * (size_t)r, hb, 0, r, c->comm, c->stream));
    }
  } else {
    NCK(g_nccl.Recv(d_boot, hb, 0, 0, c->comm, c->stream));
    NCK(g_nccl.Send(d_boot + hb * (size_t)c->rank, hb, 0, 0, c->comm, c->stream));
  }
  NCK(g_nccl.GroupEnd());
  CCK(cudaStreamSynchronize(c->stream));
  BootRec all[64];
  CCK(cudaMemcpy(all, d_boot, hb * (size_t)c->nranks, cudaMemcpyDeviceToHost));
  if (c->rank == 0) {
    for (int r = 1; r < c->nranks && ok; ++r)
      if (!all[r].ok || cudaIpcOpenMemHandle((void **)&c->peerBlock[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  } else if (ok) {
    if (!all[0].ok || cudaIpcOpenMemHandle((void **)&c->peerBlock[0], all[0].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  if (c->diag == 8 && c->rank == c->nranks - 1) ok = 0;      // B200_COMM_DIAG=8: the last rank pretends its open failed (test of the fallback)
  // verdict = AND over the ranks: gathered by rank 0, sent back
  int *d_ok = reinterpret_cast<int *>(d_boot);
  CCK(cudaMemcpy(d_ok + c->rank, &ok, sizeof(int), cudaMemcpyHostToDevice));
  if (c->rank == 0) {
    NCK(g_nccl.GroupStart());
    for (int r = 1; r < c->nranks; ++r) NCK(g_nccl.Recv(d_ok + r, sizeof(int), 0, r, c->comm, c->stream));
    NCK(g_nccl.GroupEnd());
    CCK(cudaStreamSynchronize(c->stream));
    int oks[64];
    CCK(cudaMemcpy(oks, d_ok, sizeof(int) * (size_t)c->nranks, cudaMemcpyDeviceToHost));
    for (int r = 0; r < c->nranks; ++r) ok &= (oks[r] != 0);
    CCK(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    NCK(g_nccl.GroupStart());
    for (int r = 1; r < c->nranks; ++r) NCK(g_nccl.Send(d_ok, sizeof(int), 0, r, c->comm, c->stream));
    NCK(g_nccl.GroupEnd());
    CCK(cudaStreamSynchronize(c->stream));
  } else {
    NCK(g_nccl.Send(d_ok + c->rank, sizeof(int), 0, 0, c->comm, c->stream));
    CCK(cudaStreamSynchronize(c->stream));
    NCK(g_nccl.Recv(d_ok, sizeof(int), 0, 0, c->comm, c->stream));
    CCK(cudaStreamSynchronize(c->stream));
    CCK(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
  }
  cudaFree(d_boot);
  if (ok) {
    if (c->rank == 0) {
      char *q = c->d_block;
      for (int s = 0; s < 2; ++s) { c->d_layerColor[s] = reinterpret_cast<b200_vec4u *>(q); q += colBytes; }
      for (int s = 0; s < 2; ++s) { c->d_layerDepth[s] = reinterpret_cast<float *>(q); q += depBytes; }
      c->d_ready = reinterpret_cast<unsigned *>(q);
    } else {
      c->d_free = reinterpret_cast<unsigned *>(c->d_block);
    }
    return B200_OK;
  }
  // fallback: the nccl transport on every rank
  for (int r = 0; r < 64; ++r) if (c->peerBlock[r]) { cudaIpcCloseMemHandle(c->peerBlock[r]); c->peerBlock[r] = nullptr; }
  cudaFree(c->d_block); c->d_block = nullptr;
  cudaGetLastError();
  c->push = false;
  if (c->rank == 0)
    for (int s = 0; s < 2; ++s) {
      CCK(cudaMalloc(&c->d_layerColor[s], sizeof(b200_vec4u) * c->pixels * (size_t)(c->nranks - 1)));
      CCK(cudaMalloc(&c->d_layerDepth[s], sizeof(float) * c->pixels * (size_t)(c->nranks - 1)));
    }
  return B200_OK;
}

b200_status b200_comm_create(int device, int nranks, int rank, const char id[B200_COMM_ID_BYTES], int img_w, int img_h, b200_comm **out) {
  if (out) *out = nullptr;
  if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks || img_w <= 0 || img_h <= 0) {
    snprintf(g_commErr, sizeof(g_commErr), "invalid communicator configuration");
    return B200_ERR_INVALID;
  }
  b200_comm *c = new b200_comm();
  memset(c, 0, sizeof(*c));
  c->device = device; c->nranks = nranks; c->rank = rank; c->pixels = (size_t)img_w * img_h;
  const b200_status st = comm_create_body(c, id);
  if (st != B200_OK) { snprintf(g_commErr, sizeof(g_commErr), "%s", c->err); b200_comm_destroy(c); return st; }
  c->mu = new std::mutex(); c->cv = new std::condition_variable(); c->queue = new std::deque<b200_comm::Work>();
  c->worker = new std::thread(comm_worker, c);
  *out = c;
  return B200_OK;
}

const char *b200_comm_last_error(const b200_comm *c) { return c ? c->err : g_commErr; }

// the stream operations of one hand-over (communicator's thread)
static b200_status enqueue_exchange(b200_comm *c, const b200_comm::Work &w) {
  const int slot = w.slot;
  const unsigned seq = w.seq;
  const b200_vec4u *d_color = w.d_color; const float *d_depth = w.d_depth;
  b200_vec4u *d_out_color = w.d_out_color; float *d_out_depth = w.d_out_depth;
  const int32_t *tints = w.hasTints ? w.tints : nullptr;
  const float dim_factor = w.dim_factor, tint_strength = w.tint_strength;
  CCK(cudaStreamWaitEvent(c->stream, c->evReady[slot], 0));     // the renders of this frame (recorded by submit on the engine's stream)
  const size_t n = c->pixels;
  if (c->push) {
    const size_t layers = (size_t)(c->nranks - 1), colBytes = sizeof(b200_vec4u) * n * layers, depBytes = sizeof(float) * n * layers;
    if (c->rank != 0) {
      // rank 0 has composited what this slot carried two frames ago (its free flag says so), then: two copy-engine transfers
      // into rank 0's layer buffers, then the ready flag
      if (seq > 1 && g_waitValue32((CUstream)c->stream, (CUdeviceptr)(c->d_free + slot), seq - 1, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
        snprintf(c->err, sizeof(c->err), "cuStreamWaitValue32 failed"); return B200_ERR_CUDA;
      }
      char *remote = c->peerBlock[0];
      b200_vec4u *rc = reinterpret_cast<b200_vec4u *>(remote + (size_t)slot * colBytes) + (size_t)(c->rank - 1) * n;
      float *rd = reinterpret_cast<float *>(remote + 2 * colBytes + (size_t)slot * depBytes) + (size_t)(c->rank - 1) * n;
      CCK(cudaMemcpyAsync(rc, d_color, n * sizeof(b200_vec4u), cudaMemcpyDefault, c->stream));
      CCK(cudaMemcpyAsync(rd, d_depth, n * sizeof(float), cudaMemcpyDefault, c->stream));
      FlagList f; f.n = 1;
      f.p[0] = reinterpret_cast<unsigned *>(remote + 2 * (colBytes + depBytes)) + (size_t)(c->rank - 1) * 2 + slot;
      k_raise_flags<<<1, 32, 0, c->stream>>>(f, seq);
      CCK(cudaGetLastError());
    } else {
      for (int r = 1; r < c->nranks; ++r)
        if (g_waitValue32((CUstream)c->stream, (CUdeviceptr)(c->d_ready + (size_t)(r - 1) * 2 + slot), seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
          snprintf(c->err, sizeof(c->err), "cuStreamWaitValue32 failed"); return B200_ERR_CUDA;
        }
    }
  } else if (c->rank != 0) {
    NCK(g_nccl.GroupStart());
    NCK(g_nccl.Send(d_color, n * sizeof(b200_vec4u), 0 /* ncclInt8 */, 0, c->comm, c->stream));
    NCK(g_nccl.Send(d_depth, n * sizeof(float), 0, 0, c->comm, c->stream));
    NCK(g_nccl.GroupEnd());
  } else if (c->nranks > 1) {
    NCK(g_nccl.GroupStart());
    for (int r = 1; r < c->nranks; ++r) {
      NCK(g_nccl.Recv(c->d_layerColor[slot] + (size_t)(r - 1) * n, n * sizeof(b200_vec4u), 0, r, c->comm, c->stream));
      NCK(g_nccl.Recv(c->d_layerDepth[slot] + (size_t)(r - 1) * n, n * sizeof(float), 0, r, c->comm, c->stream));
    }
    NCK(g_nccl.GroupEnd());
  }
  if (c->rank == 0) {
    // CompositeInstances (InstanceReconstructor.cpp:932-987): the background render dimmed, every instance layer z-composited
    // on top in rank order — on the communicator's stream, behind the layers' arrival
    b200_instance_layer layersArr[64];
    const int nl = c->nranks - 1 < 64 ? c->nranks - 1 : 64;
    for (int r = 0; r < nl; ++r) {
      layersArr[r].d_color = c->d_layerColor[slot] + (size_t)r * n;
      layersArr[r].d_depth = c->d_layerDepth[slot] + (size_t)r * n;
      for (int k = 0; k < 4; ++k) layersArr[r].tint[k] = tints ? tints[4 * r + k] : 0;
    }
    // rank 0's own render is the background: read in place by the composite kernel, never copied
    if (c->diag != 1) launch_composite_layers(w.e, d_out_color, d_out_depth, (int)n, layersArr, nl, dim_factor >= 0.0f, dim_factor, tint_strength, c->stream,
                                             d_color, d_depth);
    CCK(cudaGetLastError());
    if (c->push) {
      FlagList f; f.n = nl;
      for (int r = 1; r <= nl; ++r) f.p[r - 1] = reinterpret_cast<unsigned *>(c->peerBlock[r]) + slot;
      k_raise_flags<<<1, 64, 0, c->stream>>>(f, seq);      // the senders may reuse the slot
      CCK(cudaGetLastError());
    }
  }
  CCK(cudaEventRecord(c->evDone[slot], c->stream));
  return B200_OK;
}

static void comm_worker(b200_comm *c) {
  cudaSetDevice(c->device);
  for (;;) {
    b200_comm::Work w;
    {
      std::unique_lock<std::mutex> lk(*c->mu);
      c->cv->wait(lk, [&] { return c->stop || !c->queue->empty(); });
      if (c->queue->empty()) return;        // stop requested and nothing left
      w = c->queue->front();
      c->queue->pop_front();
    }
    b200_status st = B200_OK;
    if (c->workerStatus == B200_OK) st = enqueue_exchange(c, w);
    {
      std::lock_guard<std::mutex> lk(*c->mu);
      if (st != B200_OK && c->workerStatus == B200_OK) c->workerStatus = st;
      c->enqueued[w.slot] = w.seq;
    }
    c->cv->notify_all();
  }
}

// the slot's latest hand-over has been turned into stream operations (its evDone record is in place); false: the worker failed
static bool wait_enqueued(b200_comm *c, int slot) {
  std::unique_lock<std::mutex> lk(*c->mu);
  c->cv->wait(lk, [&] { return c->enqueued[slot] == c->seq[slot]; });
  return c->workerStatus == B200_OK;
}

// Per frame, every rank. d_color / d_depth: this volume's renders (w*h), written by work already enqueued on the engine's
// stream. Rank 0 additionally names the composite's destination and the compositing parameters; tints[r-1] is rank r's tint.
b200_status b200_gather_composite_submit(b200_comm *c, b200_engine *e, const b200_vec4u *d_color, const float *d_depth,
                                         b200_vec4u *d_out_color, float *d_out_depth, const int32_t *tints, float dim_factor,
                                         float tint_strength, int slot) {
  if (!c || !e || !d_color || !d_depth || slot < 0 || slot > 1) return B200_ERR_INVALID;
  if (c->rank == 0 && (!d_out_color || !d_out_depth)) { snprintf(c->err, sizeof(c->err), "rank 0 needs the composite's destination"); return B200_ERR_INVALID; }
  if (c->busy[slot]) { snprintf(c->err, sizeof(c->err), "slot %d resubmitted before b200_gather_composite_wait / _release", slot); return B200_ERR_INVALID; }
  if (c->workerStatus != B200_OK) return c->workerStatus;
  CCK(cudaSetDevice(c->device));
  CCK(cudaEventRecord(c->evReady[slot], e->stream));          // the renders of this frame
  b200_comm::Work w;
  w.slot = slot; w.d_color = d_color; w.d_depth = d_depth; w.d_out_color = d_out_color; w.d_out_depth = d_out_depth;
  w.hasTints = tints != nullptr;
  if (tints) memcpy(w.tints, tints, sizeof(int32_t) * 4 * (size_t)(c->nranks - 1 < 64 ? c->nranks - 1 : 64));
  w.dim_factor = dim_factor; w.tint_strength = tint_strength; w.e = e;
  {
    std::lock_guard<std::mutex> lk(*c->mu);
    w.seq = ++c->seq[slot];
    c->queue->push_back(w);
  }
  c->cv->notify_all();
  c->busy[slot] = true;
  return B200_OK;
}

// The engine's stream waits (on the device, not the host) until slot's exchange has drained: call before enqueueing work
// that overwrites the render buffers handed over with that slot.
b200_status b200_gather_composite_release(b200_comm *c, b200_engine *e, int slot) {
  if (!c || !e || slot < 0 || slot > 1) return B200_ERR_INVALID;
  if (!c->busy[slot]) return B200_OK;
  if (!wait_enqueued(c, slot)) return c->workerStatus;
  CCK(cudaSetDevice(c->device));
  if (c->diag != 4) CCK(cudaStreamWaitEvent(e->stream, c->evDone[slot], 0));      // (diag 4: measurement aid — the buffers may be overwritten early)
  c->busy[slot] = false;
  return B200_OK;
}

// Host-blocking: slot's exchange (and, on rank 0, the composite) has finished.
b200_status b200_gather_composite_wait(b200_comm *c, int slot) {
  if (!c || slot < 0 || slot > 1) return B200_ERR_INVALID;
  if (!c->busy[slot]) return B200_OK;
  if (!wait_enqueued(c, slot)) return c->workerStatus;
  CCK(cudaSetDevice(c->device));
  CCK(cudaEventSynchronize(c->evDone[slot]));
  CCK(cudaGetLastError());
  c->busy[slot] = false;
  return B200_OK;
}

}  // extern "C"
