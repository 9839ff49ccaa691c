// dist.cu -- data-parallel plumbing of the hot path: NCCL communicator, halo exchange of the mat-vec input
// overlapped with the interior rows, allreduce of the Krylov inner products.
//
// New work: the reference has no multi-GPU code at all (partition.c is dead, no MPI/NCCL).  The
// decomposition follows the ownership notion of reference src/partition.c:16-77 (nodal partition -> element halo):
// every node is owned by one rank, a rank assembles every element touching an owned node (ghost elements are
// recomputed, so assembly needs NO communication and stays deterministic), rows are owned, the only exchange steps
// are (1) the ghost entries of the SpMV input, (2) the sums of the inner products.
//
// NCCL is resolved with dlopen so that single-GPU users carry no NCCL dependency; with torch in the process the
// already-loaded libnccl.so.2 is reused.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace dfb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return DFB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return DFB_ERR_ARG; }
#define SYM(field, name)                                              \
  *(void**)(&g_nccl.field) = dlsym(h, name);                          \
  if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return DFB_ERR_ARG; }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.handle = h;
  return DFB_OK;
}

#define DFB_NCCL(expr)                                                                              \
  do {                                                                                              \
    ncclResult_t _r = (expr);                                                                       \
    if (_r != ncclSuccess) {                                                                        \
      set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r));          \
      return DFB_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

// gather (u,p) of the listed nodes: buf[4*t + 0..2] = x[3*node + 0..2], buf[4*t+3] = x[poff + node]
__global__ void k_halo_pack(int n, const int* __restrict__ nodes, const f64* __restrict__ x, size_t poff,
                            f64* __restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int nd = nodes[t];
  const f64* xu = x + (size_t)nd * 3;
  f64* b = buf + (size_t)t * 4;
  b[0] = xu[0]; b[1] = xu[1]; b[2] = xu[2]; b[3] = x[poff + nd];
}

__global__ void k_halo_unpack(int n, const int* __restrict__ nodes, const f64* __restrict__ buf, f64* __restrict__ x,
                              size_t poff) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int nd = nodes[t];
  const f64* b = buf + (size_t)t * 4;
  f64* xu = x + (size_t)nd * 3;
  xu[0] = b[0]; xu[1] = b[1]; xu[2] = b[2]; x[poff + nd] = b[3];
}

// the same for a vector in the solver's interleaved layout x[4 node + c]
__global__ void k_halo_pack_aos(int n, const int* __restrict__ nodes, const f64* __restrict__ x, f64* __restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  reinterpret_cast<double2*>(buf)[t] = reinterpret_cast<const double2*>(x)[(size_t)nodes[t >> 1] * 2 + (t & 1)];
}
__global__ void k_halo_unpack_aos(int n, const int* __restrict__ nodes, const f64* __restrict__ buf, f64* __restrict__ x) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  reinterpret_cast<double2*>(x)[(size_t)nodes[t >> 1] * 2 + (t & 1)] = reinterpret_cast<const double2*>(buf)[t];
}

}  // namespace dfb

using namespace dfb;

struct dfb_comm {
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  int n_local = 0;
  std::vector<int> nbr, send_off, recv_off;
  int *d_send_nodes = nullptr, *d_recv_nodes = nullptr;
  f64 *d_send_buf = nullptr, *d_recv_buf = nullptr;
  int n_send = 0, n_recv = 0;
  // peer-memory mode
  void* shared = nullptr;               // my region: mailbox words, then z
  size_t shared_bytes = 0;
  void* peer_base[P2P_MAXR] = {nullptr};
  int* d_remote_nodes = nullptr;
  int *d_tgt_ptr = nullptr, *d_tgt_q = nullptr, *d_tgt_rid = nullptr;
  f64** d_tgt_addr = nullptr;
  f64* d_nrm_part = nullptr;
  unsigned* d_push_ctr = nullptr;
  unsigned* d_err = nullptr;            // raised by a kernel whose bounded wait ran out (common.cuh p2p_give_up)
  unsigned long long seq = 0, hseq = 0; // sequence counters of the fused collectives: ONE owner per communicator
  unsigned long long* d_seq_base = nullptr;   // their values at the start of the running solve, on the device
  P2PView* d_view = nullptr;
  P2PView h_view;
  P2PHandle handle;
  bool p2p_ready = false;
};

extern "C" {

int dfb_comm_unique_id(void* id128) {
  DFB_CHECK(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  DFB_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
  return DFB_OK;
}

int dfb_comm_create(dfb_comm** out, int rank, int nranks, const void* id128) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks || !id128) { set_error("dfb_comm_create: bad argument"); return DFB_ERR_ARG; }
  DFB_CHECK(load_nccl());
  dfb_comm* c = new dfb_comm();
  c->rank = rank; c->nranks = nranks;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete c; return DFB_ERR_CUDA; }
  DFB_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  DFB_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  *out = c;
  return DFB_OK;
}

void dfb_comm_destroy(dfb_comm* c) {
  if (!c) return;
  cudaFree(c->d_send_nodes); cudaFree(c->d_recv_nodes); cudaFree(c->d_send_buf); cudaFree(c->d_recv_buf);
  for (int r = 0; r < c->nranks && r < P2P_MAXR; r++)
    if (c->peer_base[r] && r != c->rank) cudaIpcCloseMemHandle(c->peer_base[r]);
  cudaFree(c->d_tgt_ptr); cudaFree(c->d_tgt_q); cudaFree(c->d_tgt_rid); cudaFree(c->d_tgt_addr); cudaFree(c->d_nrm_part);
  cudaFree(c->shared); cudaFree(c->d_remote_nodes); cudaFree(c->d_push_ctr); cudaFree(c->d_err); cudaFree(c->d_seq_base); cudaFree(c->d_view);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
}

int dfb_comm_set_halo(dfb_comm* c, int n_local, int nn, const int* nbr, const int* send_off, const int* send_nodes,
                      const int* recv_off, const int* recv_nodes) {
  if (!c || n_local <= 0 || nn < 0) { set_error("dfb_comm_set_halo: bad argument"); return DFB_ERR_ARG; }
  c->n_local = n_local;
  c->nbr.assign(nbr, nbr + nn);
  c->send_off.assign(send_off, send_off + nn + 1);
  c->recv_off.assign(recv_off, recv_off + nn + 1);
  c->n_send = nn ? send_off[nn] : 0;
  c->n_recv = nn ? recv_off[nn] : 0;
  cudaFree(c->d_send_nodes); cudaFree(c->d_recv_nodes); cudaFree(c->d_send_buf); cudaFree(c->d_recv_buf);
  c->d_send_nodes = c->d_recv_nodes = nullptr; c->d_send_buf = c->d_recv_buf = nullptr;
  if (c->n_send) {
    DFB_CUDA(cudaMalloc(&c->d_send_nodes, sizeof(int) * (size_t)c->n_send));
    DFB_CUDA(cudaMalloc(&c->d_send_buf, sizeof(f64) * 4 * (size_t)c->n_send));
    DFB_CUDA(cudaMemcpy(c->d_send_nodes, send_nodes, sizeof(int) * (size_t)c->n_send, cudaMemcpyHostToDevice));
  }
  if (c->n_recv) {
    DFB_CUDA(cudaMalloc(&c->d_recv_nodes, sizeof(int) * (size_t)c->n_recv));
    DFB_CUDA(cudaMalloc(&c->d_recv_buf, sizeof(f64) * 4 * (size_t)c->n_recv));
    DFB_CUDA(cudaMemcpy(c->d_recv_nodes, recv_nodes, sizeof(int) * (size_t)c->n_recv, cudaMemcpyHostToDevice));
  }
  return DFB_OK;
}

int dfb_comm_allreduce(double* d_buf, int count, void* stream, void* user) {
  dfb_comm* c = reinterpret_cast<dfb_comm*>(user);
  if (!c || !d_buf || count <= 0) { set_error("dfb_comm_allreduce: bad argument"); return DFB_ERR_ARG; }
  if (c->nranks == 1) return DFB_OK;
  DFB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, c->comm, as_stream(stream)));
  count_launch();
  return DFB_OK;
}

// pack -> grouped send/recv -> unpack on the communicator's own stream, ordered after everything already
// enqueued on `stream`; the caller keeps computing on `stream` until dfb_comm_halo_end.
static int halo_begin_impl(double* d_x, void* stream, void* user, bool aos) {
  dfb_comm* c = reinterpret_cast<dfb_comm*>(user);
  if (!c || !d_x) { set_error("dfb_comm_halo_begin: bad argument"); return DFB_ERR_ARG; }
  if (c->nranks == 1 || c->nbr.empty()) return DFB_OK;
  cudaStream_t st = as_stream(stream), cs = c->comm_stream;
  const size_t poff = (size_t)3 * c->n_local;
  DFB_CUDA(cudaEventRecord(c->ev_ready, st));
  DFB_CUDA(cudaStreamWaitEvent(cs, c->ev_ready, 0));
  if (c->n_send) {
    if (aos) k_halo_pack_aos<<<ceil_div(2 * (i64)c->n_send, 256), 256, 0, cs>>>(c->n_send, c->d_send_nodes, d_x, c->d_send_buf);
    else k_halo_pack<<<ceil_div(c->n_send, 256), 256, 0, cs>>>(c->n_send, c->d_send_nodes, d_x, poff, c->d_send_buf);
    DFB_LAUNCH_CHECK();
  }
  DFB_NCCL(g_nccl.GroupStart());
  for (size_t q = 0; q < c->nbr.size(); q++) {
    const int ns = c->send_off[q + 1] - c->send_off[q], nr = c->recv_off[q + 1] - c->recv_off[q];
    if (ns) DFB_NCCL(g_nccl.Send(c->d_send_buf + (size_t)4 * c->send_off[q], (size_t)4 * ns, ncclDouble, c->nbr[q], c->comm, cs));
    if (nr) DFB_NCCL(g_nccl.Recv(c->d_recv_buf + (size_t)4 * c->recv_off[q], (size_t)4 * nr, ncclDouble, c->nbr[q], c->comm, cs));
  }
  DFB_NCCL(g_nccl.GroupEnd());
  count_launch();
  if (c->n_recv) {
    if (aos) k_halo_unpack_aos<<<ceil_div(2 * (i64)c->n_recv, 256), 256, 0, cs>>>(c->n_recv, c->d_recv_nodes, c->d_recv_buf, d_x);
    else k_halo_unpack<<<ceil_div(c->n_recv, 256), 256, 0, cs>>>(c->n_recv, c->d_recv_nodes, c->d_recv_buf, d_x, poff);
    DFB_LAUNCH_CHECK();
  }
  DFB_CUDA(cudaEventRecord(c->ev_done, cs));
  return DFB_OK;
}
int dfb_comm_halo_begin(double* d_x, void* stream, void* user) { return halo_begin_impl(d_x, stream, user, false); }
int dfb_comm_halo_begin_aos(double* d_x4, void* stream, void* user) { return halo_begin_impl(d_x4, stream, user, true); }

int dfb_comm_halo_end(double* d_x, void* stream, void* user) {
  dfb_comm* c = reinterpret_cast<dfb_comm*>(user);
  (void)d_x;
  if (!c) { set_error("dfb_comm_halo_end: bad argument"); return DFB_ERR_ARG; }
  if (c->nranks == 1 || c->nbr.empty()) return DFB_OK;
  DFB_CUDA(cudaStreamWaitEvent(as_stream(stream), c->ev_done, 0));
  return DFB_OK;
}

int dfb_comm_p2p_alloc(dfb_comm* c, void* handle64) {
  if (!c || !handle64 || c->n_local <= 0) { set_error("dfb_comm_p2p_alloc: bad argument (call dfb_comm_set_halo first)"); return DFB_ERR_ARG; }
  if (c->nranks > P2P_MAXR) { set_error("dfb_comm_p2p_alloc: at most %d ranks", P2P_MAXR); return DFB_ERR_ARG; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!c->shared) {
    c->shared_bytes = sizeof(unsigned long long) * P2P_MBOX_WORDS + sizeof(f64) * 6 * (size_t)c->n_local;
    DFB_CUDA(cudaMalloc(&c->shared, c->shared_bytes));
    DFB_CUDA(cudaMemset(c->shared, 0, c->shared_bytes));
  }
  cudaIpcMemHandle_t h;
  DFB_CUDA(cudaIpcGetMemHandle(&h, c->shared));
  memcpy(handle64, &h, 64);
  return DFB_OK;
}

int dfb_comm_p2p_connect(dfb_comm* c, const void* handles, const int* h_remote_nodes, const int* h_nbr_num_local) {
  if (!c || !handles || !c->shared || (c->n_send && !h_remote_nodes) || (!c->nbr.empty() && !h_nbr_num_local)) { set_error("dfb_comm_p2p_connect: bad argument"); return DFB_ERR_ARG; }
  if ((int)c->nbr.size() > P2P_MAXR) { set_error("dfb_comm_p2p_connect: too many neighbours"); return DFB_ERR_ARG; }
  const char* hb = static_cast<const char*>(handles);
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank) { c->peer_base[r] = c->shared; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)64 * r, 64);
    DFB_CUDA(cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  P2PView& v = c->h_view;
  memset(&v, 0, sizeof(v));
  v.rank = c->rank; v.nranks = c->nranks;
  for (int r = 0; r < c->nranks; r++) {
    v.mbox_peer[r] = static_cast<unsigned long long*>(c->peer_base[r]);
    v.z_peer[r] = reinterpret_cast<f64*>(static_cast<unsigned long long*>(c->peer_base[r]) + P2P_MBOX_WORDS);
  }
  v.mbox_local = v.mbox_peer[c->rank];
  v.z_local = v.z_peer[c->rank];
  v.n_nbr = (int)c->nbr.size();
  for (int q = 0; q < v.n_nbr; q++) {
    v.nbr[q] = c->nbr[q];
    v.send_off[q] = c->send_off[q];
    v.nbr_poff[q] = (unsigned long long)3 * (unsigned long long)h_nbr_num_local[q];
  }
  v.send_off[v.n_nbr] = c->n_send;
  if (c->n_send) {
    cudaFree(c->d_remote_nodes);
    DFB_CUDA(cudaMalloc(&c->d_remote_nodes, sizeof(int) * (size_t)c->n_send));
    DFB_CUDA(cudaMemcpy(c->d_remote_nodes, h_remote_nodes, sizeof(int) * (size_t)c->n_send, cudaMemcpyHostToDevice));
  }
  v.send_nodes = c->d_send_nodes;
  v.remote_nodes = c->d_remote_nodes;
  // invert the send lists: per boundary-owned node the (neighbour slot, remote id) targets
  v.tgt_base = 0; v.tgt_n = 0;
  if (c->n_send) {
    std::vector<int> sn((size_t)c->n_send);
    DFB_CUDA(cudaMemcpy(sn.data(), c->d_send_nodes, sizeof(int) * (size_t)c->n_send, cudaMemcpyDeviceToHost));
    int lo = sn[0], hi = sn[0];
    for (int x : sn) { lo = x < lo ? x : lo; hi = x > hi ? x : hi; }
    const int n = hi - lo + 1;
    std::vector<int> ptr((size_t)n + 1, 0), tq((size_t)c->n_send), trid((size_t)c->n_send);
    for (int x : sn) ptr[(size_t)(x - lo) + 1]++;
    for (int i = 0; i < n; i++) ptr[(size_t)i + 1] += ptr[i];
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int q = 0; q < v.n_nbr; q++)
      for (int t = c->send_off[q]; t < c->send_off[q + 1]; t++) {
        const int pos = fill[(size_t)(sn[t] - lo)]++;
        tq[pos] = q;
        trid[pos] = h_remote_nodes[t];
      }
    std::vector<f64*> taddr((size_t)c->n_send);
    for (int t = 0; t < c->n_send; t++) taddr[(size_t)t] = v.z_peer[v.nbr[tq[(size_t)t]]] + (size_t)trid[(size_t)t] * 4;
    cudaFree(c->d_tgt_ptr); cudaFree(c->d_tgt_q); cudaFree(c->d_tgt_rid); cudaFree(c->d_tgt_addr);
    DFB_CUDA(cudaMalloc(&c->d_tgt_addr, sizeof(f64*) * (size_t)c->n_send));
    DFB_CUDA(cudaMemcpy(c->d_tgt_addr, taddr.data(), sizeof(f64*) * (size_t)c->n_send, cudaMemcpyHostToDevice));
    DFB_CUDA(cudaMalloc(&c->d_tgt_ptr, sizeof(int) * ((size_t)n + 1)));
    DFB_CUDA(cudaMalloc(&c->d_tgt_q, sizeof(int) * (size_t)c->n_send));
    DFB_CUDA(cudaMalloc(&c->d_tgt_rid, sizeof(int) * (size_t)c->n_send));
    DFB_CUDA(cudaMemcpy(c->d_tgt_ptr, ptr.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    DFB_CUDA(cudaMemcpy(c->d_tgt_q, tq.data(), sizeof(int) * (size_t)c->n_send, cudaMemcpyHostToDevice));
    DFB_CUDA(cudaMemcpy(c->d_tgt_rid, trid.data(), sizeof(int) * (size_t)c->n_send, cudaMemcpyHostToDevice));
    v.tgt_base = lo; v.tgt_n = n;
    v.tgt_ptr = c->d_tgt_ptr; v.tgt_q = c->d_tgt_q; v.tgt_rid = c->d_tgt_rid; v.tgt_addr = c->d_tgt_addr;
  }
  if (!c->d_nrm_part) {
    DFB_CUDA(cudaMalloc(&c->d_nrm_part, sizeof(f64)));
    DFB_CUDA(cudaMemset(c->d_nrm_part, 0, sizeof(f64)));
  }
  v.nrm_part = c->d_nrm_part;
  if (!c->d_push_ctr) {
    DFB_CUDA(cudaMalloc(&c->d_push_ctr, sizeof(unsigned)));
    DFB_CUDA(cudaMemset(c->d_push_ctr, 0, sizeof(unsigned)));
  }
  v.push_ctr = c->d_push_ctr;
  if (!c->d_err) {
    DFB_CUDA(cudaMalloc(&c->d_err, sizeof(unsigned)));
    DFB_CUDA(cudaMemset(c->d_err, 0, sizeof(unsigned)));
  }
  v.err = c->d_err;
  if (!c->d_seq_base) {
    DFB_CUDA(cudaMalloc(&c->d_seq_base, 2 * sizeof(unsigned long long)));
    DFB_CUDA(cudaMemset(c->d_seq_base, 0, 2 * sizeof(unsigned long long)));
  }
  v.seq_base = c->d_seq_base;
  {   // wait budget of one poll loop (default 10 s; DFB_P2P_TIMEOUT_MS overrides, tests use a short one)
    const char* e = getenv("DFB_P2P_TIMEOUT_MS");
    const double ms = e ? atof(e) : 10000.0;
    v.timeout_ns = (unsigned long long)((ms > 0.0 ? ms : 10000.0) * 1e6);
  }
  if (!c->d_view) DFB_CUDA(cudaMalloc(&c->d_view, sizeof(P2PView)));
  DFB_CUDA(cudaMemcpy(c->d_view, &v, sizeof(P2PView), cudaMemcpyHostToDevice));
  c->p2p_ready = true;
  return DFB_OK;
}

// host copy of the view (solve.cu reads sizes from it and passes the device copy, stored right behind it, to kernels)
const void* dfb_comm_p2p_view(dfb_comm* c) {
  if (!c || !c->p2p_ready) return nullptr;
  c->handle.host = c->h_view;
  c->handle.dev = c->d_view;
  c->handle.seq = &c->seq;
  c->handle.hseq = &c->hseq;
  c->handle.d_err = c->d_err;
  c->handle.d_seq_base = c->d_seq_base;
  return &c->handle;
}

int dfb_comm_halo(dfb_comm* c, double* d_x, void* stream) {
  DFB_CHECK(dfb_comm_halo_begin(d_x, stream, c));
  return dfb_comm_halo_end(d_x, stream, c);
}

}  // extern "C"
