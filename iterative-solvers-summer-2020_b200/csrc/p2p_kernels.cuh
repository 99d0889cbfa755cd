// Peer-memory collectives of the slab decomposition over NVLink / NVSwitch (one process per GPU, library-owned
// buffers mapped into every peer with CUDA IPC):
//   * p2p_halo_kernel     -- halo exchange by direct stores: each rank writes its first / last two rows straight into
//                            its ring neighbours' halo buffers, then raises their arrival flags and waits for its own;
//   * p2p_allreduce_kernel -- one-shot all-reduce of <= 64 fp64 scalars: every rank stores its partials into every
//                            peer's mailbox, raises a flag, waits for all flags and sums the P rows in a fixed pairwise tree
//                            (identical bits on every rank, so the replicated Hessenberg/Givens state cannot diverge).
// Both replace an NCCL call (send/recv pair, ncclAllReduce) whose cost at these sizes (256 KiB, <= 50 doubles) is pure
// launch + protocol latency.  Flags carry monotonically increasing epochs; payload buffers are double-buffered by
// epoch parity, which is sufficient because a rank can run at most one exchange ahead of a peer (completing exchange
// e needs the peer's contribution to e, which the peer sends only after it has consumed exchange e-1).  The all-reduce
// mailboxes are FOUR deep (epoch & 3): kernels of a speculatively enqueued Arnoldi step that find JS_STOP set skip their
// exchange on every rank alike, so up to two consecutive epochs may never be used and two live exchanges can be three
// epochs apart.
#pragma once
#include "cuda_common.cuh"

namespace jfnk {

constexpr int kP2PMaxRanks = 8;
constexpr int kP2PMaxScalars = 64;
constexpr int kP2PHaloSlots = 3; // linearisation point, operand, generic (as CudaOps::halo_ptrs)
constexpr unsigned long long kP2PTimeoutNs = 120ull * 1000000000ull; // ranks may start tens of seconds apart

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= epoch ; a peer that never arrives sets *err instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch, int* err) {
  if (ld_acquire_sys(flag) >= epoch) return;
  const unsigned long long t0 = global_timer_ns();
  for (unsigned spin = 1;; ++spin) {
    if (ld_acquire_sys(flag) >= epoch) return;
    if ((spin & 1023u) == 0 && global_timer_ns() - t0 > kP2PTimeoutNs) { *err = 1; return; }
  }
}

struct P2PHaloArgs {
  const double* src_first;  // my first two rows  -> previous rank's bottom halo
  const double* src_last;   // my last two rows   -> next rank's top halo
  double* dst_prev_bot;     // peer-mapped destination buffers (this epoch's parity)
  double* dst_next_top;
  unsigned long long* flag_prev_bot; // in the previous rank's block: "bottom halo arrived"
  unsigned long long* flag_next_top; // in the next rank's block:     "top halo arrived"
  const unsigned long long* my_flag_top; // raised by the previous rank
  const unsigned long long* my_flag_bot; // raised by the next rank
  unsigned long long epoch;
  size_t count;             // doubles per message (2 nx)
  unsigned* ticket;         // zero between launches
  int* err;
};

__global__ void __launch_bounds__(256) p2p_halo_kernel(P2PHaloArgs A) {
  __shared__ bool is_last;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.count; i += stride) {
    A.dst_prev_bot[i] = A.src_first[i];
    A.dst_next_top[i] = A.src_last[i];
  }
  // One system-scope fence per CTA (by the thread that takes the ticket, after the block barrier), not one per thread:
  // a MEMBAR.SYS from each of 8192 threads costs ~20 us.  The fences are cumulative over the writes ordered before
  // them by the barrier / the ticket, so the rows of every CTA are visible system-wide before the flags are raised.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    is_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) {
    *A.ticket = 0u;
    __threadfence_system();
    st_release_sys(A.flag_prev_bot, A.epoch);
    st_release_sys(A.flag_next_top, A.epoch);
  }
  if (threadIdx.x < 2) wait_flag(threadIdx.x == 0 ? A.my_flag_top : A.my_flag_bot, A.epoch, A.err);
}

struct P2PReduceArgs {
  double* S;                 // local scalar arena: S[off .. off+cnt) reduced in place
  int off, cnt, op;          // op 0: sum, 1: max, 2: entry 0 sum and the others max (the three norms of a residual pass)
  int rank, nranks;
  double* mailbox_peer[kP2PMaxRanks];             // peer q's mailbox of this parity: row `rank` is mine to write
  unsigned long long* flags_peer[kP2PMaxRanks];   // peer q's arrival flags: entry `rank` is mine to raise
  const double* my_mailbox;                       // [nranks][kP2PMaxScalars]
  const unsigned long long* my_flags;             // [nranks]
  unsigned long long epoch;
  int givens_j, givens_taken, givens_rerun;       // >= 0: run the Hessenberg/Givens step of column j afterwards
  int skippable;                                  // part of the Arnoldi loop: dropped while S[JS_STOP] is set
  int* err;
};

// The all-reduce proper, run by ONE CTA of at least kP2PMaxScalars threads (all of its threads must call it).  Used by the
// stand-alone kernel below and, fused, by the CTA that finalises the local sums of a multi-dot / Gram-Schmidt update
// (blas_kernels.cuh): the collective then costs no launch of its own.
__device__ __forceinline__ void p2p_allreduce_block(const P2PReduceArgs& A) {
  const int i = threadIdx.x;
  __syncthreads(); // the local values S[off..] were written by other threads of this CTA
  if (i < A.cnt) {
    const double v = A.S[A.off + i];
    for (int q = 0; q < A.nranks; ++q) A.mailbox_peer[q][(size_t)A.rank * kP2PMaxScalars + i] = v;
  }
  __syncthreads();
  if (i < A.nranks) {
    __threadfence_system(); // cumulative over the CTA's mailbox stores (ordered before it by the barrier)
    st_release_sys(A.flags_peer[i] + A.rank, A.epoch);
    wait_flag(A.my_flags + i, A.epoch, A.err);
  }
  __syncthreads();
  if (i < A.cnt) {
    // sums: the fixed pairwise tree of the rank-count-independent reductions (cuda_common.cuh), the same bits on every rank
    double v[kP2PMaxRanks];
#pragma unroll
    for (int r = 0; r < kP2PMaxRanks; ++r) v[r] = (r < A.nranks) ? __ldcg(A.my_mailbox + (size_t)r * kP2PMaxScalars + i) : 0.0;
    double acc;
    if (A.op == 1 || (A.op == 2 && i > 0)) {
      acc = v[0];
      for (int r = 1; r < A.nranks; ++r) acc = fmax(acc, v[r]);
    } else acc = tree_sum(v, A.nranks);
    A.S[A.off + i] = acc;
  }
  if (A.givens_j >= 0) {
    __syncthreads();
    if (i == 0) hess_givens_step(A.S, A.givens_j, A.givens_taken, A.givens_rerun);
  }
}

__global__ void __launch_bounds__(kP2PMaxScalars) p2p_allreduce_kernel(P2PReduceArgs A) {
  if (A.skippable && A.S[JS_STOP] != 0.0) return; // speculative Arnoldi work after the process stopped (every rank agrees)
  p2p_allreduce_block(A);
}

} // namespace jfnk
