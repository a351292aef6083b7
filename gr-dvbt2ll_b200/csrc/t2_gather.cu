// Ordered multi-GPU reassembly of finished T2 frames on one GPU (C ABI: dvbt2ll_gather_* in include/dvbt2ll_cuda.h).
//
// The chain shards by channel / T2-frame run with no data-path collective; the only exchange is the reassembly of the
// ranks' output into ONE ordered stream on the root GPU -- the single sink of the reference flowgraph
// (apps/vv009-4kshort.grc:1696-1697).  Design:
//
//   * the root owns a ring of n_slots "steps", each slot_bytes long, laid out in final stream order;
//   * the root's own chain writes straight into the slot (no copy); every other rank produces into a local buffer
//     (double-buffered) and PUSHES it into the root slot at its ordered offset with one peer copy over NVLink
//     (copy engine, on a side stream, so it overlaps the rank's next step);
//   * completion and back-pressure are device-side counters, no host handshake: after its copy a rank stores step+1
//     into arrived[rank] in the root's control block (a one-thread kernel storing to peer memory); the root's consumer
//     stream waits on the counters with stream memory operations (cuStreamWaitValue32, no spinning kernel); when the
//     consumer is done with a slot the root stores step+1 into every rank's released counter, which the rank's side
//     stream waits on before it overwrites that slot n_slots steps later.
//
// Ranks may be processes (one per GPU, torchrun: buffers are shared through CUDA IPC handles carried in the
// connect blobs) or handles of one process driving several devices (raw pointers + peer access).
#include "../../include/dvbt2ll_cuda.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace t2k { void count_extra_launch(); }

namespace {

int gfail(int code, const std::string &msg);

#define GCK(call)                                                                                 \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return gfail(DVBT2LL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

__global__ void k_signal(volatile uint32_t *flag, uint32_t value)
{
  __threadfence_system();
  *flag = value;
  __threadfence_system();
}

struct Blob {                  // one per rank, DVBT2LL_GATHER_BLOB_BYTES
  uint32_t magic, rank;
  int32_t device;
  int64_t pid;
  uint64_t ctrl_ptr, ring_ptr;                 // raw device pointers (valid inside the exporting process)
  cudaIpcMemHandle_t ctrl_ipc, ring_ipc;       // ring only meaningful for the root
  uint8_t pad[DVBT2LL_GATHER_BLOB_BYTES - 4 - 4 - 4 - 4 - 8 - 8 - 8 - 2 * sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(Blob) == DVBT2LL_GATHER_BLOB_BYTES, "blob size is part of the ABI");

struct Ctrl {                  // device-resident control block of a rank
  uint32_t arrived[64];        // root: arrived[r] = (last step of rank r that has fully landed) + 1
  uint32_t released;           // every rank: (last step whose root slot the consumer has released) + 1
  uint32_t pad[63];
};

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

} // namespace

struct dvbt2ll_gather {
  int rank, world, root, device, n_slots;
  size_t slot_bytes;
  bool connected;
  Ctrl *ctrl;                         // own control block (device memory of `device`)
  uint8_t *ring;                      // root: the ring (own memory); others: the root's ring as mapped here
  Ctrl *root_ctrl;                    // the root's control block as mapped here
  std::vector<Ctrl *> peer_ctrl;      // root: every rank's control block as mapped here
  std::vector<void *> ipc_opened;
  uint8_t *local[2];                  // non-root: local production buffers
  size_t local_bytes;
  cudaStream_t side;                  // push stream
  cudaEvent_t produced, pushed[2], released_ev[8];
  long long steps_released;
  WaitValue32Fn wait32;
  dvbt2ll_gather() : connected(false), ctrl(0), ring(0), root_ctrl(0), local_bytes(0), side(0), produced(0), steps_released(0), wait32(0)
  {
    local[0] = local[1] = 0; pushed[0] = pushed[1] = 0;
    for (int i = 0; i < 8; i++) released_ev[i] = 0;
  }
};

namespace {

thread_local std::string g_gerr;
int gfail(int code, const std::string &msg) { g_gerr = msg; return code; }

int wait_geq(dvbt2ll_gather *g, cudaStream_t s, const uint32_t *addr, uint32_t value)
{
  CUresult r = g->wait32((CUstream)s, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) return gfail(DVBT2LL_ERR_CUDA, "cuStreamWaitValue32 failed (" + std::to_string((int)r) + ")");
  return 0;
}

} // namespace

extern "C" {

const char *dvbt2ll_gather_last_error(void) { return g_gerr.c_str(); }

dvbt2ll_gather *dvbt2ll_gather_create(int rank, int world, int root, int device, size_t slot_bytes, size_t local_bytes, int n_slots)
{
  if (world < 1 || world > 64 || rank < 0 || rank >= world || root < 0 || root >= world || n_slots < 1 || n_slots > 8 || slot_bytes == 0) {
    gfail(DVBT2LL_ERR_INVALID, "gather: bad arguments");
    return 0;
  }
  dvbt2ll_gather *g = new dvbt2ll_gather();
  g->rank = rank; g->world = world; g->root = root; g->device = device; g->n_slots = n_slots; g->slot_bytes = slot_bytes;
  g->local_bytes = local_bytes;
  g->peer_ctrl.assign(world, (Ctrl *)0);
  bool ok = cudaSetDevice(device) == cudaSuccess;
  ok = ok && cudaMalloc((void **)&g->ctrl, sizeof(Ctrl)) == cudaSuccess;
  ok = ok && cudaMemset(g->ctrl, 0, sizeof(Ctrl)) == cudaSuccess;
  if (ok && rank == root) ok = cudaMalloc((void **)&g->ring, slot_bytes * (size_t)n_slots) == cudaSuccess;
  if (ok && rank != root)
    for (int i = 0; i < 2 && ok; i++) ok = cudaMalloc((void **)&g->local[i], local_bytes ? local_bytes : 16) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->produced, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < 2 && ok; i++) ok = cudaEventCreateWithFlags(&g->pushed[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < n_slots && ok; i++) ok = cudaEventCreateWithFlags(&g->released_ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (ok) {
    void *fn = 0;
    cudaDriverEntryPointQueryResult qr;
    ok = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn && qr == cudaDriverEntryPointSuccess;
    g->wait32 = (WaitValue32Fn)fn;
  }
  ok = ok && cudaDeviceSynchronize() == cudaSuccess;
  if (!ok) {
    gfail(DVBT2LL_ERR_CUDA, std::string("gather: CUDA setup failed: ") + cudaGetErrorString(cudaGetLastError()));
    dvbt2ll_gather_destroy(g);
    return 0;
  }
  return g;
}

int dvbt2ll_gather_export(dvbt2ll_gather *g, void *blob, size_t cap)
{
  if (!g || !blob || cap < sizeof(Blob)) return gfail(DVBT2LL_ERR_INVALID, "gather: export buffer too small");
  GCK(cudaSetDevice(g->device));
  Blob b;
  std::memset(&b, 0, sizeof(b));
  b.magic = 0x54324741u; b.rank = (uint32_t)g->rank; b.device = g->device; b.pid = (int64_t)getpid();
  b.ctrl_ptr = (uint64_t)(uintptr_t)g->ctrl; b.ring_ptr = (uint64_t)(uintptr_t)g->ring;
  GCK(cudaIpcGetMemHandle(&b.ctrl_ipc, g->ctrl));
  if (g->rank == g->root) GCK(cudaIpcGetMemHandle(&b.ring_ipc, g->ring));
  std::memcpy(blob, &b, sizeof(b));
  return (int)sizeof(Blob);
}

// blobs: the world's export blobs concatenated in rank order
int dvbt2ll_gather_connect(dvbt2ll_gather *g, const void *blobs, size_t bytes)
{
  if (!g || !blobs || bytes < sizeof(Blob) * (size_t)g->world) return gfail(DVBT2LL_ERR_INVALID, "gather: connect needs one blob per rank");
  GCK(cudaSetDevice(g->device));
  const Blob *B = reinterpret_cast<const Blob *>(blobs);
  for (int r = 0; r < g->world; r++)
    if (B[r].magic != 0x54324741u || (int)B[r].rank != r) return gfail(DVBT2LL_ERR_INVALID, "gather: blobs are not in rank order");
  const int64_t me = (int64_t)getpid();
  // map `ipc` / `raw` of rank r into this process
  auto map = [&](const Blob &b, const cudaIpcMemHandle_t &ipc, uint64_t raw, void **out) -> int {
    if (b.pid == me) {
      if (b.device != g->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return gfail(DVBT2LL_ERR_CUDA, std::string("gather: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      *out = (void *)(uintptr_t)raw;
      return 0;
    }
    GCK(cudaIpcOpenMemHandle(out, ipc, cudaIpcMemLazyEnablePeerAccess));
    g->ipc_opened.push_back(*out);
    return 0;
  };
  int rc;
  if (g->rank == g->root) {
    g->root_ctrl = g->ctrl;
    for (int r = 0; r < g->world; r++) {
      if (r == g->root) { g->peer_ctrl[r] = g->ctrl; continue; }
      void *p = 0;
      if ((rc = map(B[r], B[r].ctrl_ipc, B[r].ctrl_ptr, &p))) return rc;
      g->peer_ctrl[r] = (Ctrl *)p;
    }
  }
  else {
    void *p = 0;
    if ((rc = map(B[g->root], B[g->root].ctrl_ipc, B[g->root].ctrl_ptr, &p))) return rc;
    g->root_ctrl = (Ctrl *)p;
    if ((rc = map(B[g->root], B[g->root].ring_ipc, B[g->root].ring_ptr, &p))) return rc;
    g->ring = (uint8_t *)p;
  }
  g->connected = true;
  return 0;
}

// Where this rank produces step `step` (offset = byte position of its part inside the ordered slot), with
// `producer` made to wait until that buffer may be overwritten.
int dvbt2ll_gather_acquire(dvbt2ll_gather *g, long long step, size_t offset, void *producer, void **ptr)
{
  if (!g || !g->connected || !ptr || step < 0) return gfail(DVBT2LL_ERR_INVALID, "gather: not connected");
  GCK(cudaSetDevice(g->device));
  cudaStream_t ps = (cudaStream_t)producer;
  if (g->rank == g->root) {
    if (offset >= g->slot_bytes) return gfail(DVBT2LL_ERR_INVALID, "gather: offset outside the slot");
    // the slot was last used by step - n_slots: its release was recorded on the consumer stream
    if (step >= g->n_slots) {
      if (g->steps_released < step - g->n_slots + 1) return gfail(DVBT2LL_ERR_INVALID, "gather: root acquires a slot that was never released");
      GCK(cudaStreamWaitEvent(ps, g->released_ev[step % g->n_slots], 0));
    }
    *ptr = g->ring + (size_t)(step % g->n_slots) * g->slot_bytes + offset;
  }
  else {
    if (step >= 2) GCK(cudaStreamWaitEvent(ps, g->pushed[step & 1], 0));     // push of step - 2 has drained the buffer
    *ptr = g->local[step & 1];
  }
  return 0;
}

// After `producer` has finished this rank's part of step `step`: ship it.  Non-root: one peer copy into the root slot at
// `offset` on the side stream, then the arrival counter.  Root: only the arrival counter (it produced in place).
int dvbt2ll_gather_push(dvbt2ll_gather *g, long long step, size_t offset, size_t bytes, void *producer)
{
  if (!g || !g->connected || step < 0) return gfail(DVBT2LL_ERR_INVALID, "gather: not connected");
  if (offset + bytes > g->slot_bytes) return gfail(DVBT2LL_ERR_INVALID, "gather: part does not fit the slot");
  GCK(cudaSetDevice(g->device));
  cudaStream_t ps = (cudaStream_t)producer;
  if (g->rank == g->root) {
    k_signal<<<1, 1, 0, ps>>>(&g->ctrl->arrived[g->rank], (uint32_t)(step + 1));
    t2k::count_extra_launch();
    GCK(cudaGetLastError());
    return 0;
  }
  if (bytes > g->local_bytes) return gfail(DVBT2LL_ERR_INVALID, "gather: part larger than the local buffer");
  GCK(cudaEventRecord(g->produced, ps));
  GCK(cudaStreamWaitEvent(g->side, g->produced, 0));
  // back-pressure: the root slot is free once step - n_slots has been released by the consumer
  if (step >= g->n_slots) {
    int rc = wait_geq(g, g->side, &g->ctrl->released, (uint32_t)(step - g->n_slots + 1));
    if (rc) return rc;
  }
  uint8_t *dst = g->ring + (size_t)(step % g->n_slots) * g->slot_bytes + offset;
  GCK(cudaMemcpyAsync(dst, g->local[step & 1], bytes, cudaMemcpyDeviceToDevice, g->side));
  k_signal<<<1, 1, 0, g->side>>>(&g->root_ctrl->arrived[g->rank], (uint32_t)(step + 1));
  t2k::count_extra_launch();
  GCK(cudaGetLastError());
  GCK(cudaEventRecord(g->pushed[step & 1], g->side));
  return 0;
}

// Root: `consumer` waits until every rank's part of `step` has landed; *slot = the ordered output of the step.
int dvbt2ll_gather_wait(dvbt2ll_gather *g, long long step, void *consumer, void **slot)
{
  if (!g || !g->connected || g->rank != g->root || step < 0) return gfail(DVBT2LL_ERR_INVALID, "gather: wait is a root call");
  GCK(cudaSetDevice(g->device));
  for (int r = 0; r < g->world; r++) {
    int rc = wait_geq(g, (cudaStream_t)consumer, &g->ctrl->arrived[r], (uint32_t)(step + 1));
    if (rc) return rc;
  }
  if (slot) *slot = g->ring + (size_t)(step % g->n_slots) * g->slot_bytes;
  return 0;
}

// Root: the consumer (everything queued on `consumer` so far) is done with the slot of `step`.
int dvbt2ll_gather_release(dvbt2ll_gather *g, long long step, void *consumer)
{
  if (!g || !g->connected || g->rank != g->root || step < 0) return gfail(DVBT2LL_ERR_INVALID, "gather: release is a root call");
  GCK(cudaSetDevice(g->device));
  cudaStream_t cs = (cudaStream_t)consumer;
  for (int r = 0; r < g->world; r++) {
    if (r == g->root) continue;
    k_signal<<<1, 1, 0, cs>>>(&g->peer_ctrl[r]->released, (uint32_t)(step + 1));
    t2k::count_extra_launch();
  }
  GCK(cudaGetLastError());
  GCK(cudaEventRecord(g->released_ev[step % g->n_slots], cs));
  if (step + 1 > g->steps_released) g->steps_released = step + 1;
  return 0;
}

// Stream that carries this rank's pushes (time the end of a run on it: cudaEventRecord after the last push).
void *dvbt2ll_gather_side_stream(dvbt2ll_gather *g) { return g ? (void *)g->side : 0; }

void dvbt2ll_gather_destroy(dvbt2ll_gather *g)
{
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (size_t i = 0; i < g->ipc_opened.size(); i++) cudaIpcCloseMemHandle(g->ipc_opened[i]);
  if (g->rank == g->root && g->ring) cudaFree(g->ring);
  for (int i = 0; i < 2; i++) if (g->local[i]) cudaFree(g->local[i]);
  if (g->ctrl) cudaFree(g->ctrl);
  if (g->side) cudaStreamDestroy(g->side);
  if (g->produced) cudaEventDestroy(g->produced);
  for (int i = 0; i < 2; i++) if (g->pushed[i]) cudaEventDestroy(g->pushed[i]);
  for (int i = 0; i < 8; i++) if (g->released_ev[i]) cudaEventDestroy(g->released_ev[i]);
  cudaGetLastError();
  delete g;
}

} // extern "C"
