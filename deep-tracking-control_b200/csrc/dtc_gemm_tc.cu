// Tensor-core GEMMs for the MLP stack on sm_100a: tcgen05.mma kind::tf32 with TMEM accumulators, operands staged by TMA
// into 128-byte-swizzled shared memory, 3-4 stage mbarrier pipeline, warp-specialised (TMA producer / MMA issuer /
// TMEM allocator / 8 epilogue warps), persistent over the tile list.  Two kernels share the epilogue:
//   k_gemm_tc   one CTA per 128x128 tile (rollout shapes, narrow layers),
//   k_gemm_tc2  a CTA pair (cta_group::2) per 256x128 tile, B halves shared across the TPC (the 24 576-row update shapes).
//
// fp32 accuracy (the 1e-5 parity bar) comes from error-compensated 3xTF32: the tensor core TRUNCATES fp32 operands to
// TF32 (measured, tools/tc_probe.cu), so with  x_lo = rn_tf32(x - trunc_tf32(x))  kept next to every operand,
//     A B^T  ~=  A.B + A_lo.B + A.B_lo          (three MMAs per k-step into the same TMEM accumulator)
// leaves a relative error of ~2^-21 per product - the same order as fp32 summation-order noise.
// The TMEM accumulator add rounds toward zero (measured: ~0.25 ulp systematic error per accumulation step), so the two
// small correction products go to their OWN accumulator (one truncating add per k-step on the main one instead of three),
// the two are summed in fp32 in the epilogue, and reductions longer than TC_MAX_KB k-blocks are split and summed in fp32.
//
// Same contract as the SIMT family (dtc_gemm.cu): C[m,n] = epi(sum_k A(m,k) B(n,k)), each operand k-contiguous
// ("K-major": TMA box 32k x 128 rows, SWIZZLE_128B) or k-strided ("MN-major": four boxes 32mn x 32k, SWIZZLE_128B with
// 32-byte atoms - the only MN-major layout kind::tf32 accepts).  Out-of-range rows / k-tails are zero-filled by TMA.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "dtc_gemm.cuh"

#define TC_BM 128
#define TC_BN 128
#define TC_BK 32
#define TC_STAGES 3
#define TC_TMEM_COLS 512
#define TC_MAX_KB 32     // k-blocks (of 32) accumulated in TMEM before the fp32 split-K sum takes over
#define TC_TILE_BYTES (128 * 32 * 4)
#define TC_STAGE_BYTES (4 * TC_TILE_BYTES)
#define TC_STG_BYTES (8 * 32 * 32 * 4)  // epilogue staging: 8 warps x 32 rows x 32 floats (XOR-swizzled)
#define TC_THREADS 512              // warp 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..11 epilogue, 12..15 operand splitters
#define TC_SMEM_BYTES (TC_STAGES * TC_STAGE_BYTES + TC_STG_BYTES + 1024)

struct TcParams {
  int M, N, K;
  int kb_per_split, nkb;
  float* C; float* C_lo; int ldc;
  const float* bias; const float* act_src; int ld_act; int epi; int accumulate;
  float* ws;
  float* colsum_part;  // [ceil(M/32)][round4(N)] column sums per 32-row block of the final output, or NULL
  float* colsum_out; int colsum_n;  // or: column sums of C[:, :colsum_n] ADDED into colsum_out (pre-zeroed) with L2 reductions
  int splits, has_alo, has_blo;
  int split_a, split_b;  // 1: that operand's TF32 companion tile is computed in shared memory by the splitter warps (no *_lo array in HBM)
  int neff;   // 1: the MMA of a ragged / narrow n-tile covers only the live columns rounded up to the instruction granularity
  int splitk_atomic;  // TMA-store epilogue, splits > 1: partial tiles are ADDED into C by the TMA unit (no workspace, no reduce kernel)
  int pdl_late;   // single-CTA kernel: griddepcontrol.launch_dependents after the last MMA instead of at kernel start (DTC_PDL=3)
  int lo_direct;  // TMA-store epilogue: 1 = the companion output leaves from registers (env DTC_TC_LO=direct)
  int direct; // epilogue variant: 1 = registers -> global without the shared-memory transpose (env DTC_TC_EPI=direct|staged)
  int debug;  // timing experiments only (env DTC_TC_DEBUG): 1 = epilogue skips its stores, 2 = producer stops loading after the first ring fill
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident while its stream predecessor is still draining; griddepcontrol.wait blocks until that predecessor has completed and
// its memory is visible, launch_dependents lets this kernel's own successor do the same.  Both are no-ops for a normal launch.
__device__ __forceinline__ void tc_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tTC_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni TC_WAIT_DONE;\n\tbra.uni TC_WAIT_LOOP;\n\tTC_WAIT_DONE:\n\t}" ::"r"(tc_smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_tma_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"((uint64_t)map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout <<61.
//   K-major : rows of 128 B (32 k), 8-row swizzle atoms 1024 B apart (SBO), LBO 16; one MMA (K = 8) advances 32 B inside the row
//   MN-major: blocks of 32 mn (4096 B apart = LBO) of 32 k-rows x 128 B, 4-row swizzle atoms 512 B apart (SBO); one MMA = 8 k-rows
// columns the MMAs of the n-tile starting at n0 must produce: all 128, or - for the ragged last tile of the 693 / 588 / 532-wide
// layers and for the 35..64-wide ones - the live columns rounded up to `gran` (16 for cta_group::1, 32 for a CTA pair, whose two
// halves each supply N/2 rows of B)
__device__ __forceinline__ int tc_n_eff(const TcParams& p, int n0, int gran) {
  if (!p.neff) return TC_BN;
  const int live = min(TC_BN, p.N - n0);
  return min(TC_BN, (live + gran - 1) / gran * gran);
}
// one elected lane of a converged warp (CUTLASS elect_one_sync): unlike `lane == 0`, ptxas knows the guarded region has a single
// active thread and issues UTMALDG / UTCHMMA / UTCBAR straight from uniform registers instead of wrapping each in an ELECT loop
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
// descriptor halves: the low word carries the start address (and LBO), the high word is constant per operand layout, so the
// per-k8 / per-stage descriptor update is one 32-bit add
// Warp-specialised register budget: the kernels run 512 threads (128 registers each at launch); the control warpgroup (TMA producer, MMA
// issuer, TMEM allocator) and the splitter warpgroup hand registers back, the two epilogue warpgroups - which hold a 32x32 fp32 chunk plus
// its correction / mask block per warp - take them: 128 x 64 + 128 x 64 + 256 x 184 = 63 488 <= 65 536.
#define TC_REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 64;")
#define TC_REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 184;")
template <int MAJ> __device__ __forceinline__ uint32_t tc_desc_hi() {
  return MAJ == 0 ? (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29)) : (uint32_t)((512u >> 4) | (1u << 14) | (1u << 29));
}
template <int MAJ> __device__ __forceinline__ uint32_t tc_desc_lo(uint32_t tile) {
  return MAJ == 0 ? (((tile >> 4) & 0x3FFFu) | ((16u >> 4) << 16)) : (((tile >> 4) & 0x3FFFu) | ((4096u >> 4) << 16));
}
template <int MAJ> __device__ __forceinline__ constexpr uint32_t tc_desc_k8_step() { return MAJ == 0 ? (32u >> 4) : (1024u >> 4); }
template <int CG>
__device__ __forceinline__ void tc_mma_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tc_epi(float v, int epi, float bias, float src) {
  switch (epi) {
    case EPI_BIAS: return v + bias;
    case EPI_BIAS_RELU: v += bias; return v > 0.f ? v : 0.f;
    case EPI_BIAS_ELU: v += bias; return v > 0.f ? v : expm1f(v);
    case EPI_DRELU: return src > 0.f ? v : 0.f;
    case EPI_DELU: return src > 0.f ? v : v * (src + 1.0f);
    default: return v;
  }
}

// Epilogue of one 128-row x 128-column accumulator set by EIGHT warps: warp w drains TMEM lanes 32 (w & 3) .. +31 (the quadrant its
// tcgen05.ld may touch) and the column half (w - 4) >> 2, in two chunks of 32 columns.  A chunk goes TMEM -> registers (main +
// correction accumulators summed in fp32) -> a 32x32 XOR-swizzled staging tile -> global, so that every global access is a
// coalesced 128-byte row segment; the act'(x) operand of a chunk is fetched with eight independent loads before any of it is used.
// The accumulator set is handed back to the MMA issuer (acc_empty: 8 arrivals per CTA) as soon as this warp's TMEM reads are done.
// ---- TMA-store epilogue (default).  What paces these kernels on the learner's shapes is the LENGTH OF THE EPILOGUE'S INSTRUCTION
// CHAIN, not bytes: ncu on the staged epilogue below shows ~800 dependent instructions per 32x32 chunk and warp (shared-memory
// transpose, per-row address arithmetic, 16 predicated STG.128, a run-time switch per element) at ~16 cycles each with only two
// epilogue warps per scheduler - 26 k cycles per 256x128 tile against 20.5 k cycles of MMAs at K = 512 (epilogue-bound: 178 of the
// 232 TFLOP/s the same kernel reaches with its stores compiled out), and ~9 us of pure epilogue on the 64..256-wide layers whose
// K loop is a few k-blocks.  Here every lane keeps its own row: bias / activation / derivative mask / fan-in accumulate are applied in
// registers, the 32x32 chunk goes to shared memory once (128-byte rows, 16-byte pieces XOR-ed with row % 8 = the layout
// CU_TENSOR_MAP_SWIZZLE_128B expects, conflict-free for the row-wise STS.128) and ONE cp.async.bulk.tensor store (UTMASTG) per chunk
// and output moves it to global memory; the TMA unit clips ragged right / bottom edges.  ~3x fewer instructions per chunk.
__device__ __forceinline__ void tc_tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
// shared -> global with an fp32 ADD at the destination (L2 reduction): split-K partial tiles accumulate straight into C
__device__ __forceinline__ void tc_tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tc_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// A warp owns NBUF staging chunks of 4 KB used round-robin, one TMA store per use: before a chunk is rewritten at most NBUF - 1 of the
// lane-0 store groups may still be reading shared memory.
template <int NBUF> __device__ __forceinline__ void tc_stg_acquire(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
  __syncwarp();
}
// the 32 values of this lane's row -> staging chunk (128-byte rows, 16-byte pieces XOR-ed with row % 8)
__device__ __forceinline__ void tc_stage_rows(const float (&v)[32], float* stg, int lane) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4)
    *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
}
// mode 0: 2-D store, 1: 3-D store into plane z (split-K workspace), 2: 2-D reduce-add (split-K partials summed in L2)
__device__ __forceinline__ void tc_store_chunk(float* stg, int lane, bool issue, const CUtensorMap* map, int mode, int gn, int gm_box, int z) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
  __syncwarp();
  if (lane == 0 && issue) {
    if (mode == 1) tc_tma_store_3d(map, tc_smem_u32(stg), gn, gm_box, z);
    else if (mode == 2) tc_tma_reduce_add_2d(map, tc_smem_u32(stg), gn, gm_box);
    else tc_tma_store_2d(map, tc_smem_u32(stg), gn, gm_box);
  }
  if (lane == 0) tc_bulk_commit();  // an (empty) group even when nothing was issued keeps the round-robin count in step
}
// A 32-row x 32-column block of a row-major global matrix, read COALESCED (lane -> 16-byte piece lane & 7 of rows (lane >> 3) + 4 i:
// eight lanes cover one 128-byte row segment), for the epilogue's per-row arithmetic.  Rows >= M and pieces past round4(N) read as 0.
struct TcAux { float4 r[8]; };
__device__ __forceinline__ TcAux tc_aux_load(const float* base, int ld, int gm_box, int gn, int M, int n4, int lane) {
  TcAux a;
  const int cb = lane & 7, r0 = lane >> 3;
  const bool col_ok = gn + 4 * cb < n4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = gm_box + r0 + 4 * i;
    a.r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && gm < M) a.r[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)gm * ld + gn) + cb);
  }
  return a;
}
// coalesced registers -> staging chunk -> this lane's row (the chunk must have been acquired; it is free again on return)
__device__ __forceinline__ void tc_aux_to_rows(const TcAux& a, float* stg, int lane, float (&out)[32]) {
  const int cb = lane & 7, r0 = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + 4 * i;
    *reinterpret_cast<float4*>(stg + r * 32 + ((cb ^ (r & 7)) << 2)) = a.r[i];
  }
  __syncwarp();
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 t = *reinterpret_cast<const float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2));
    out[j4 * 4] = t.x; out[j4 * 4 + 1] = t.y; out[j4 * 4 + 2] = t.z; out[j4 * 4 + 3] = t.w;
  }
  __syncwarp();
}
// Waits for the tile's accumulator (acc_full) itself, so that the derivative-mask block of the first chunk is already in flight.
template <int BN, int NBUF>
__device__ __forceinline__ void tc_epilogue_tma(const TcParams& p, const CUtensorMap* mapC, const CUtensorMap* mapClo, uint32_t tmem, int buf,
                                                int warp, int lane, int m0, int n0, int z, float* stg_base, int& stg_turn,
                                                uint64_t* acc_full, uint32_t acc_full_parity, uint32_t acc_empty_addr, bool cluster_arrive) {
  const int q = warp & 3, half = (warp - 4) >> 2;
  const bool partial = p.splits > 1;
  const int n4 = (p.N + 3) & ~3;
  const int epi = partial ? EPI_STORE : p.epi;
  const bool has_bias = epi >= EPI_BIAS && epi <= EPI_BIAS_ELU, has_src = epi == EPI_DRELU || epi == EPI_DELU;
  const bool accum = !partial && p.accumulate;
  const bool has_corr = (p.has_alo || p.has_blo) && !(p.debug & 16);
  const bool want_lo = !partial && p.C_lo != nullptr;
  const bool colsum = (p.colsum_part || p.colsum_out) && !partial;
  int nchunks = (p.N - n0 + 31) / 32;  // live 32-column chunks of this tile
  nchunks = nchunks > BN / 32 ? BN / 32 : nchunks;
  const int ch_first = (BN / 64) * half;
  const int ch_last = min(ch_first + BN / 64 - 1, nchunks - 1);  // last chunk this warp reads (< ch_first: none)
  const int gm_box = m0 + q * 32, gm = gm_box + lane;
  const bool row_ok = gm < p.M, box_ok = gm_box < p.M;
  TcAux src;
  if (has_src && ch_last >= ch_first) src = tc_aux_load(p.act_src, p.ld_act, gm_box, n0 + ch_first * 32, p.M, n4, lane);
  tc_mbar_wait(acc_full, acc_full_parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ch_last < ch_first) {  // nothing to read: release the accumulator set right away
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if (cluster_arrive) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
    }
    return;
  }
#pragma unroll 1
  for (int ch = ch_first; ch <= ch_last; ++ch) {
    const int c0 = ch * 32, gn = n0 + c0;
    float v[32];
    {
      uint32_t u[32], w[32];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
          "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
            "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
            "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
            "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
          : "r"(taddr));
      if (has_corr) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
            "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
              "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]),
              "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]),
              "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
            : "r"(taddr + BN));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = has_corr ? __uint_as_float(u[i]) + __uint_as_float(w[i]) : __uint_as_float(u[i]);
    }
    if (ch == ch_last) {
      // all TMEM reads of this warp are done: hand the accumulator set back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (cluster_arrive) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
      }
    }
    if (p.debug & 1) continue;
    float* stg = stg_base + (stg_turn % NBUF) * (32 * 32);
    tc_stg_acquire<NBUF>(lane);  // `stg` is free from here until the store below is issued
    // ---- epilogue function on this lane's row (columns gn .. gn + 31; pieces past round4(N) are clipped by the store)
    if (has_bias) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const int col = gn + 4 * j4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col + 3 < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));  // same address on every lane: one broadcast
        else {
          if (col < p.N) b4.x = __ldg(p.bias + col);
          if (col + 1 < p.N) b4.y = __ldg(p.bias + col + 1);
          if (col + 2 < p.N) b4.z = __ldg(p.bias + col + 2);
        }
        v[j4 * 4] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
      }
      if (epi == EPI_BIAS_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.f;
      } else if (epi == EPI_BIAS_ELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
      }
    } else if (has_src) {
      float sv[32];
      tc_aux_to_rows(src, stg, lane, sv);
      if (ch < ch_last) src = tc_aux_load(p.act_src, p.ld_act, gm_box, gn + 32, p.M, n4, lane);  // next chunk's block: in flight under this one
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float sx = sv[i], x = v[i];
        v[i] = epi == EPI_DRELU ? (sx > 0.f ? x : 0.f) : (sx > 0.f ? x : x * (sx + 1.0f));
      }
    }
    if (accum) {
      float ov[32];
      const TcAux old = tc_aux_load(p.C, p.ldc, gm_box, gn, p.M, n4, lane);
      tc_aux_to_rows(old, stg, lane, ov);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += ov[i];
    }
    if (colsum && !row_ok) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    tc_stage_rows(v, stg, lane);
    tc_store_chunk(stg, lane, box_ok, mapC, partial ? (p.splitk_atomic ? 2 : 1) : 0, gn, gm_box, z);
    ++stg_turn;
    if (colsum) {
      // column sums of the 32 staged rows, lane = column: piece (lane >> 2) ^ (r & 7) of row r - 32 distinct banks per read.  The
      // TMA unit may be reading the chunk at the same time; nothing writes it before the next acquire.
      float cs = 0.f;
#pragma unroll
      for (int r = 0; r < 32; ++r) cs += stg[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
      if (p.colsum_out) { if (box_ok && gn + lane < p.colsum_n) atomicAdd(p.colsum_out + gn + lane, cs); }
      else if (box_ok && gn + lane < n4) p.colsum_part[(size_t)(gm_box >> 5) * n4 + gn + lane] = cs;
    }
    if (want_lo) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = tf32_lo(v[i]);
      if (p.lo_direct) {
        // companion straight from registers (lane = row, 8 x 16 bytes): no second wait on the staging chunk's TMA store
        if (row_ok) {
          float* lrow = p.C_lo + (size_t)gm * p.ldc + gn;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            if (gn + 4 * j4 < n4) *(reinterpret_cast<float4*>(lrow) + j4) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        }
      } else {
        stg = stg_base + (stg_turn % NBUF) * (32 * 32);
        tc_stg_acquire<NBUF>(lane);
        tc_stage_rows(v, stg, lane);
        tc_store_chunk(stg, lane, box_ok, mapClo, 0, gn, gm_box, 0);
        ++stg_turn;
      }
    }
  }
}

template <int BN>
__device__ __forceinline__ void tc_epilogue_tile(const TcParams& p, uint32_t tmem, int buf, int warp, int lane, int m0, int n0, int z,
                                                 float* stg, uint32_t acc_empty_addr, bool cluster_arrive) {
  // BN = columns of one accumulator (128: k_gemm_tc / k_gemm_tc2, 256: k_gemm_tc3); the correction accumulator sits BN columns
  // behind the main one; each of the two warps of a lane quadrant drains BN / 2 columns in chunks of 32
  const int q = warp & 3, half = (warp - 4) >> 2;
  const bool partial = p.splits > 1;
  const int n4 = (p.N + 3) & ~3;
  const int ldo = partial ? n4 : p.ldc;
  const int epi = partial ? EPI_STORE : p.epi;
  const bool has_bias = epi >= EPI_BIAS && epi <= EPI_BIAS_ELU, has_src = epi == EPI_DRELU || epi == EPI_DELU;
  const bool accum = !partial && p.accumulate;
  const bool has_corr = p.has_alo || p.has_blo;
  float* const out = partial ? p.ws + (size_t)z * p.M * n4 : p.C;
  float* const out_lo = (!partial && p.C_lo) ? p.C_lo : nullptr;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  int nchunks = (p.N - n0 + 31) / 32;  // live 32-column chunks of this tile
  nchunks = nchunks > BN / 32 ? BN / 32 : nchunks;
  const int ch_first = (BN / 64) * half;
  const int ch_last = min(ch_first + BN / 64 - 1, nchunks - 1);  // last chunk this warp reads (< ch_first: none)
  if (ch_last < ch_first) {  // nothing to read: release the accumulator set right away
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if (cluster_arrive) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
    }
    return;
  }
#pragma unroll 1
  for (int ch = ch_first; ch <= ch_last; ++ch) {
    const int c0 = ch * 32;
    float v[32];
    {
      uint32_t u[32], w[32];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
          "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
            "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
            "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
            "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
          : "r"(taddr));
      if (has_corr) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
            "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
              "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]),
              "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]),
              "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
            : "r"(taddr + BN));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = has_corr ? __uint_as_float(u[i]) + __uint_as_float(w[i]) : __uint_as_float(u[i]);
    }
    if (ch == ch_last) {
      // all TMEM reads of this warp are done: hand the accumulator set back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (cluster_arrive) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty_addr) : "memory");
      }
    }
    if (p.debug & 1) continue;
    if (p.direct && n0 + c0 + 31 < p.N) {
      // ---- direct variant: lane = row, its 32 consecutive columns straight from registers to global memory.  No shared-memory
      // round trip (8 STS.128 + 8 LDS.128 per lane and chunk on a pipe the MMA operand reads already saturate); every lane writes
      // whole 128-byte row segments, two 16-byte pieces of a sector in back-to-back instructions.
      const int gm = m0 + q * 32 + lane, gn = n0 + c0;
      const bool row_ok = gm < p.M;
      if (has_bias) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn) + j4);  // same address on every lane: one broadcast
          v[j4 * 4] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
        }
        if (epi == EPI_BIAS_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.f;
        } else if (epi == EPI_BIAS_ELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
        }
      } else if (has_src) {
        const float4* sp = reinterpret_cast<const float4*>(p.act_src + (size_t)(row_ok ? gm : 0) * p.ld_act + gn);
        float4 s4[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) s4[j4] = __ldg(sp + j4);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float sv[4] = {s4[j4].x, s4[j4].y, s4[j4].z, s4[j4].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float sx = sv[c], x = v[j4 * 4 + c];
            v[j4 * 4 + c] = epi == EPI_DRELU ? (sx > 0.f ? x : 0.f) : (sx > 0.f ? x : x * (sx + 1.0f));
          }
        }
      }
      float* orow = out + (size_t)(row_ok ? gm : 0) * ldo + gn;
      if (accum && row_ok) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 o4 = *(reinterpret_cast<const float4*>(orow) + j4);
          v[j4 * 4] += o4.x; v[j4 * 4 + 1] += o4.y; v[j4 * 4 + 2] += o4.z; v[j4 * 4 + 3] += o4.w;
        }
      }
      if (row_ok) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) *(reinterpret_cast<float4*>(orow) + j4) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        if (out_lo) {
          float* lrow = out_lo + (size_t)gm * ldo + gn;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *(reinterpret_cast<float4*>(lrow) + j4) = make_float4(tf32_lo(v[j4 * 4]), tf32_lo(v[j4 * 4 + 1]), tf32_lo(v[j4 * 4 + 2]), tf32_lo(v[j4 * 4 + 3]));
        }
      }
      if ((p.colsum_part || p.colsum_out) && !partial) {
        // column sums of the 32 rows: butterfly transpose-reduce (31 shuffles), lane l ends up with column l
        if (!row_ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) {
          const bool up = (lane & sft) != 0;
#pragma unroll
          for (int i = 0; i < sft; ++i) {
            const float send = up ? v[i] : v[i + sft], keep = up ? v[i + sft] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
          }
        }
        if (p.colsum_out) { if (m0 + q * 32 < p.M && gn + lane < p.colsum_n) atomicAdd(p.colsum_out + gn + lane, v[0]); }
        else if (m0 + q * 32 < p.M) p.colsum_part[(size_t)((m0 + q * 32) >> 5) * n4 + gn + lane] = v[0];
      }
      continue;
    }
    // row `lane`, 16-byte block j4 -> physical block j4 ^ (lane & 7): conflict-free for the row-wise writes and the
    // 4-rows-x-8-blocks reads below
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
      *reinterpret_cast<float4*>(stg + lane * 32 + ((j4 ^ (lane & 7)) << 2)) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
    __syncwarp();
    const int cb = lane & 7, c4 = cb * 4, gn = n0 + c0 + c4;
    const int r0 = lane >> 3, gm0 = m0 + q * 32 + r0;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of this lane's 8 rows (colsum_part)
    if (gn + 3 < p.N) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn));
      float4 s4[8], o4[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int gm = gm0 + 4 * i;
        s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        o4[i] = s4[i];
        if (gm < p.M) {
          if (has_src) s4[i] = __ldg(reinterpret_cast<const float4*>(p.act_src + (size_t)gm * p.ld_act + gn));
          if (accum) o4[i] = *reinterpret_cast<const float4*>(out + (size_t)gm * ldo + gn);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + 4 * i, gm = gm0 + 4 * i;
        if (gm >= p.M) continue;
        const float4 a4 = *reinterpret_cast<const float4*>(stg + r * 32 + ((cb ^ (r & 7)) << 2));
        float4 x;
        x.x = tc_epi(a4.x, epi, b4.x, s4[i].x) + o4[i].x; x.y = tc_epi(a4.y, epi, b4.y, s4[i].y) + o4[i].y;
        x.z = tc_epi(a4.z, epi, b4.z, s4[i].z) + o4[i].z; x.w = tc_epi(a4.w, epi, b4.w, s4[i].w) + o4[i].w;
        *reinterpret_cast<float4*>(out + (size_t)gm * ldo + gn) = x;
        if (out_lo) *reinterpret_cast<float4*>(out_lo + (size_t)gm * ldo + gn) = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
        cs.x += x.x; cs.y += x.y; cs.z += x.z; cs.w += x.w;
      }
    } else if (gn < p.N) {  // ragged right edge: scalar tail
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + 4 * i, gm = gm0 + 4 * i;
        if (gm >= p.M) continue;
        const float4 a4 = *reinterpret_cast<const float4*>(stg + r * 32 + ((cb ^ (r & 7)) << 2));
        const float x[4] = {a4.x, a4.y, a4.z, a4.w};
        float* orow = out + (size_t)gm * ldo + gn;
        for (int jj = 0; jj < 4 && gn + jj < p.N; ++jj) {
          const float bias = has_bias ? __ldg(p.bias + gn + jj) : 0.f;
          const float src = has_src ? __ldg(p.act_src + (size_t)gm * p.ld_act + gn + jj) : 0.f;
          float y = tc_epi(x[jj], epi, bias, src);
          if (accum) y += orow[jj];
          orow[jj] = y;
          if (out_lo) out_lo[(size_t)gm * ldo + gn + jj] = tf32_lo(y);
          if (jj == 0) cs.x += y; else if (jj == 1) cs.y += y; else if (jj == 2) cs.z += y; else cs.w += y;
        }
      }
    }
    __syncwarp();
    if ((p.colsum_part || p.colsum_out) && !partial) {  // warp-uniform, outside the per-lane branches above: every lane takes part in the shuffles
      // lanes with the same column block (lane & 7) hold the four row groups: fold them, lanes 0..7 store (n4 covers gn + 3)
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
      if (p.colsum_out) {
        if (lane < 8 && m0 + q * 32 < p.M) {
          if (gn < p.colsum_n) atomicAdd(p.colsum_out + gn, cs.x);
          if (gn + 1 < p.colsum_n) atomicAdd(p.colsum_out + gn + 1, cs.y);
          if (gn + 2 < p.colsum_n) atomicAdd(p.colsum_out + gn + 2, cs.z);
          if (gn + 3 < p.colsum_n) atomicAdd(p.colsum_out + gn + 3, cs.w);
        }
      } else if (lane < 8 && gn < p.N && m0 + q * 32 < p.M)
        *reinterpret_cast<float4*>(p.colsum_part + (size_t)((m0 + q * 32) >> 5) * n4 + gn) = cs;
    }
  }
}

// ---- in-SM operand split.  The 3xTF32 companion x_lo = rn_tf32(x - trunc_tf32(x)) of an ACTIVATION operand used to live in HBM
// next to x: every producer wrote 8 bytes per element and every consumer pulled 8 bytes per element through L2 and TMA - a third of a
// forward GEMM's operand stream and half of its stores, on kernels whose stores queue behind a 190 KB-deep TMA load pipeline.  With
// split_a / split_b the TMA loads only x; warps 12..15 turn each landed tile into its companion tile in shared memory (element-wise,
// so the same byte offset whatever the swizzle / major-ness), fence the generic-proxy writes for the tensor core's async-proxy reads
// and arrive on split[s], which the MMA issuer waits on instead of full[s].
__device__ __forceinline__ void tc_split_tile(const uint8_t* src, uint8_t* dst, int n16, int tid) {
#pragma unroll 4
  for (int i = tid; i < n16; i += 128) {
    const float4 x = *reinterpret_cast<const float4*>(src + (size_t)i * 16);
    *reinterpret_cast<float4*>(dst + (size_t)i * 16) = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
  }
}

// Persistent kernel: one CTA per SM walks the (split, m-tile, n-tile) list with stride gridDim.x.  TMEM holds two
// accumulator sets (main | correction, 2 x 128 columns each), so the epilogue of tile j overlaps the MMAs of tile j+1; the TMA
// producer runs ahead across tile boundaries.
template <int AMAJ, int BMAJ>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapB,
          const __grid_constant__ CUtensorMap mapBlo, const __grid_constant__ CUtensorMap mapC,
           const __grid_constant__ CUtensorMap mapClo, const TcParams p) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_STAGES], bar_empty[TC_STAGES], bar_split[TC_STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (tc_smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt_n = (p.N + TC_BN - 1) / TC_BN, nt_m = (p.M + TC_BM - 1) / TC_BM;
  const int ntiles = nt_n * nt_m * p.splits;
  if (!p.pdl_late) tc_launch_dependents();

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(&bar_full[s], 1); tc_mbar_init(&bar_empty[s], 1); tc_mbar_init(&bar_split[s], 4); }
    for (int s = 0; s < 2; ++s) { tc_mbar_init(&bar_acc_full[s], 1); tc_mbar_init(&bar_acc_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the previous kernel's tail; its
  // results are visible from here on
  tc_grid_dependency_wait();

  if (warp < 4) {
  TC_REG_DEC();
  if (warp == 0) {
   if (tc_elect_one()) {
    // ------------------------------------------------------------ TMA producer
    const uint32_t bytes = TC_TILE_BYTES * (2 + (p.has_alo && !p.split_a) + (p.has_blo && !p.split_b));
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int n0 = (t % nt_n) * TC_BN, m0 = ((t / nt_n) % nt_m) * TC_BM, z = t / (nt_n * nt_m);
      const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TC_STAGES;
        tc_mbar_wait(&bar_empty[s], ((it / TC_STAGES) & 1) ^ 1);
        tc_mbar_expect_tx(&bar_full[s], bytes);
        const uint32_t st = smem0 + s * TC_STAGE_BYTES;
        if (AMAJ == 0) {
          tc_tma_2d(st, &mapA, &bar_full[s], kb * TC_BK, m0);
          if (p.has_alo && !p.split_a) tc_tma_2d(st + TC_TILE_BYTES, &mapAlo, &bar_full[s], kb * TC_BK, m0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tc_tma_2d(st + j * 4096, &mapA, &bar_full[s], m0 + 32 * j, kb * TC_BK);
            if (p.has_alo && !p.split_a) tc_tma_2d(st + TC_TILE_BYTES + j * 4096, &mapAlo, &bar_full[s], m0 + 32 * j, kb * TC_BK);
          }
        }
        if (BMAJ == 0) {
          tc_tma_2d(st + 2 * TC_TILE_BYTES, &mapB, &bar_full[s], kb * TC_BK, n0);
          if (p.has_blo && !p.split_b) tc_tma_2d(st + 3 * TC_TILE_BYTES, &mapBlo, &bar_full[s], kb * TC_BK, n0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tc_tma_2d(st + 2 * TC_TILE_BYTES + j * 4096, &mapB, &bar_full[s], n0 + 32 * j, kb * TC_BK);
            if (p.has_blo && !p.split_b) tc_tma_2d(st + 3 * TC_TILE_BYTES + j * 4096, &mapBlo, &bar_full[s], n0 + 32 * j, kb * TC_BK);
          }
        }
      }
    }
   }
  } else if (warp == 1) {
   if (tc_elect_one()) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    // instruction descriptor: D f32 | A,B tf32 | majors | N >> 3 | M >> 4   (cute::UMMA::InstrDescriptor)
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)AMAJ << 15) | ((uint32_t)BMAJ << 16) |
                            ((uint32_t)(TC_BM >> 4) << 24);
    int it = 0, j = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
      const int z = t / (nt_n * nt_m);
      const uint32_t idesc = idesc0 | ((uint32_t)(tc_n_eff(p, (t % nt_n) * TC_BN, 16) >> 3) << 17);
      const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
      const int buf = j & 1;
      tc_mbar_wait(&bar_acc_empty[buf], ((j >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_main = tmem + (uint32_t)buf * (2 * TC_BN), d_corr = d_main + TC_BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TC_STAGES;
        tc_mbar_wait((p.split_a | p.split_b) ? &bar_split[s] : &bar_full[s], (it / TC_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem0 + s * TC_STAGE_BYTES;
        const uint32_t a0 = tc_desc_lo<AMAJ>(st), al0 = tc_desc_lo<AMAJ>(st + TC_TILE_BYTES);
        const uint32_t b0 = tc_desc_lo<BMAJ>(st + 2 * TC_TILE_BYTES), bl0 = tc_desc_lo<BMAJ>(st + 3 * TC_TILE_BYTES);
        const uint32_t ahi = tc_desc_hi<AMAJ>(), bhi = tc_desc_hi<BMAJ>();
#pragma unroll
        for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
          const uint32_t ka = k8 * tc_desc_k8_step<AMAJ>(), kb8 = k8 * tc_desc_k8_step<BMAJ>();
          const uint32_t first = (kb == kb0 && k8 == 0) ? 0u : 1u;
          tc_mma_lh<1>(d_main, a0 + ka, ahi, b0 + kb8, bhi, idesc, first);
          if (p.has_alo) tc_mma_lh<1>(d_corr, al0 + ka, ahi, b0 + kb8, bhi, idesc, first);
          if (p.has_blo) tc_mma_lh<1>(d_corr, a0 + ka, ahi, bl0 + kb8, bhi, idesc, (first || p.has_alo) ? 1u : 0u);
        }
        tc_commit(&bar_empty[s]);  // frees the stage once these MMAs have read it
      }
      tc_commit(&bar_acc_full[buf]);  // accumulator set complete
    }
    // late trigger: the dependent grid may become resident now that this CTA has issued its last MMA - its prologue (barrier
    // init, TMEM allocation, descriptor fetch) overlaps this grid's epilogue instead of spinning next to its main loop
    if (p.pdl_late) tc_launch_dependents();
   }
  }
  } else if (warp >= 12) {
    TC_REG_DEC();
    // ------------------------------------------------------------ operand splitters (4 warps), see tc_split_tile
    if (p.split_a | p.split_b) {
      uint8_t* const base = tc_smem_raw + (smem0 - tc_smem_u32(tc_smem_raw));
      const int tid = threadIdx.x - 384;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int z = t / (nt_n * nt_m);
        const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % TC_STAGES;
          tc_mbar_wait(&bar_full[s], (it / TC_STAGES) & 1);
          uint8_t* const st = base + (size_t)s * TC_STAGE_BYTES;
          if (p.split_a) tc_split_tile(st, st + TC_TILE_BYTES, TC_TILE_BYTES / 16, tid);
          if (p.split_b) tc_split_tile(st + 2 * TC_TILE_BYTES, st + 3 * TC_TILE_BYTES, TC_TILE_BYTES / 16, tid);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&bar_split[s])) : "memory");
        }
      }
    }
  } else {
    TC_REG_INC();
    // ------------------------------------------------------------ epilogue: 8 warps, TMEM -> registers -> smem -> global
    float* const stg = reinterpret_cast<float*>(tc_smem_raw + (smem0 - tc_smem_u32(tc_smem_raw)) + TC_STAGES * TC_STAGE_BYTES) + (warp - 4) * (32 * 32);
    int j = 0, stg_turn = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
      const int n0 = (t % nt_n) * TC_BN, m0 = ((t / nt_n) % nt_m) * TC_BM, z = t / (nt_n * nt_m);
      const int buf = j & 1;
      if (p.direct == 2) {
        tc_epilogue_tma<TC_BN, 1>(p, &mapC, &mapClo, tmem, buf, warp, lane, m0, n0, z, stg, stg_turn, &bar_acc_full[buf], (j >> 1) & 1,
                                  tc_smem_u32(&bar_acc_empty[buf]), false);
      } else {
        tc_mbar_wait(&bar_acc_full[buf], (j >> 1) & 1);
        tc_epilogue_tile<TC_BN>(p, tmem, buf, warp, lane, m0, n0, z, stg, tc_smem_u32(&bar_acc_empty[buf]), false);
      }
    }
    if (p.direct == 2 && lane == 0) tc_bulk_wait_read();  // the staging chunks must outlive the last stores' reads
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

// ================================================================== CTA-pair variant (cta_group::2)
// ncu on the kernel above: the tensor pipe is 57 % busy and the shared-memory pipe is the limiter - one 128x128x8 TF32 MMA
// reads 8 KB of operands (64 wavefronts) in its 64 cycles, and the TMA writes of the four operand tiles (16 KB per k8-step of
// three MMAs) take another 128 wavefronts, i.e. 320 wavefronts per 192 MMA cycles.  Two CTAs of one TPC working on a 256x128
// tile halve the B traffic: each CTA stages its 128 rows of A / A_lo and HALF (64 rows) of B / B_lo, the leader's
// tcgen05.mma.cta_group::2 (M = 256) reads each B half once for both tensor cores -> 240 wavefronts per 192 MMA cycles.
//   * cluster (2,1,1); the pair walks (split, 256-row m-tile, n-tile) with stride gridDim.x / 2;
//   * full[s]   lives in the leader: one arrival (its producer, expect_tx = the bytes of BOTH CTAs); the peer's TMA
//               (cp.async.bulk.tensor...cta_group::2) completes its bytes on the leader's barrier;
//   * empty[s], acc_full[b] live in both CTAs and are signalled by multicast tcgen05.commit from the leader;
//   * acc_empty[b] lives in the leader: 8 arrivals (4 epilogue warps x 2 CTAs, the peer's through mapa);
//   * each CTA drains its own 128 TMEM lanes with the same epilogue as above.
#define TC2_STAGES 4
#define TC2_STG_BUFS 1                     // staging chunks per epilogue warp (TMA-store epilogue: C and C_lo stores in flight together)
#define TC2_A_BYTES TC_TILE_BYTES          // 128 rows x 32 k fp32
#define TC2_B_BYTES (TC_TILE_BYTES / 2)    // 64 rows x 32 k fp32
#define TC2_STAGE_BYTES (2 * TC2_A_BYTES + 2 * TC2_B_BYTES)
#define TC2_SMEM_BYTES (TC2_STAGES * TC2_STAGE_BYTES + TC2_STG_BUFS * TC_STG_BYTES + 1024)

__device__ __forceinline__ uint32_t tc_cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t tc_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of this CTA's tile whose bytes complete on `bar_cluster` (a shared::cluster address, possibly in the peer CTA)
__device__ __forceinline__ void tc2_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1)
               : "memory");
}
// arrives (once all MMAs issued so far have completed) on the barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void tc2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <int AMAJ, int BMAJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
k_gemm_tc2(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapB,
           const __grid_constant__ CUtensorMap mapBlo, const __grid_constant__ CUtensorMap mapC,
           const __grid_constant__ CUtensorMap mapClo, const TcParams p) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC2_STAGES], bar_empty[TC2_STAGES], bar_split[TC2_STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (tc_smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc_cluster_ctarank();
  const int nt_n = (p.N + TC_BN - 1) / TC_BN, nt_m = (p.M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int ntiles = nt_n * nt_m * p.splits;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  tc_launch_dependents();

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC2_STAGES; ++s) { tc_mbar_init(&bar_full[s], 1); tc_mbar_init(&bar_empty[s], 1); tc_mbar_init(&bar_split[s], 8); }
    for (int s = 0; s < 2; ++s) { tc_mbar_init(&bar_acc_full[s], 1); tc_mbar_init(&bar_acc_empty[s], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  tc_cluster_sync();  // both CTAs' barriers are initialised before any remote arrival / multicast commit
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  tc_grid_dependency_wait();  // programmatic dependent launch (see k_gemm_tc)

  if (warp < 4) {
  TC_REG_DEC();
  if (warp == 0) {
   if (tc_elect_one()) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const bool split = (p.split_a | p.split_b) != 0;  // then each CTA's tiles complete on its OWN full[s] (its splitters wait there)
    const bool load_alo = p.has_alo && !p.split_a, load_blo = p.has_blo && !p.split_b;
    const uint32_t bytes_cta = (uint32_t)(TC2_A_BYTES * (1 + load_alo) + TC2_B_BYTES * (1 + load_blo));
    int it = 0;
    for (int t = pair; t < ntiles; t += npairs) {
      const int nt0 = (t % nt_n) * TC_BN;
      const int n0 = nt0 + (int)rank * (tc_n_eff(p, nt0, 32) / 2), m0 = ((t / nt_n) % nt_m) * (2 * TC_BM) + (int)rank * TC_BM;
      const int z = t / (nt_n * nt_m);
      const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TC2_STAGES;
        tc_mbar_wait(&bar_empty[s], ((it / TC2_STAGES) & 1) ^ 1);
        const uint32_t full = split ? tc_smem_u32(&bar_full[s]) : tc_mapa(tc_smem_u32(&bar_full[s]), 0);  // own / the leader's barrier
        if ((p.debug & 2) && it >= TC2_STAGES) {
          if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&bar_full[s])) : "memory");
          continue;
        }
        if (split) tc_mbar_expect_tx(&bar_full[s], bytes_cta);
        else if (rank == 0) tc_mbar_expect_tx(&bar_full[s], 2 * bytes_cta);
        const uint32_t st = smem0 + s * TC2_STAGE_BYTES;
        const uint32_t sA = st, sAlo = st + TC2_A_BYTES, sB = st + 2 * TC2_A_BYTES, sBlo = sB + TC2_B_BYTES;
        if (AMAJ == 0) {
          tc2_tma_2d(sA, &mapA, full, kb * TC_BK, m0);
          if (load_alo) tc2_tma_2d(sAlo, &mapAlo, full, kb * TC_BK, m0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tc2_tma_2d(sA + j * 4096, &mapA, full, m0 + 32 * j, kb * TC_BK);
            if (load_alo) tc2_tma_2d(sAlo + j * 4096, &mapAlo, full, m0 + 32 * j, kb * TC_BK);
          }
        }
        if (BMAJ == 0) {
          tc2_tma_2d(sB, &mapB, full, kb * TC_BK, n0);
          if (load_blo) tc2_tma_2d(sBlo, &mapBlo, full, kb * TC_BK, n0);
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            tc2_tma_2d(sB + j * 4096, &mapB, full, n0 + 32 * j, kb * TC_BK);
            if (load_blo) tc2_tma_2d(sBlo + j * 4096, &mapBlo, full, n0 + 32 * j, kb * TC_BK);
          }
        }
      }
    }
   }
  } else if (warp == 1 && rank == 0) {
   if (tc_elect_one()) {
    // ------------------------------------------------------------ MMA issuer (one thread of the leader CTA), M = 256
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)AMAJ << 15) | ((uint32_t)BMAJ << 16) |
                            ((uint32_t)((2 * TC_BM) >> 4) << 24);
    int it = 0, j = 0;
    for (int t = pair; t < ntiles; t += npairs, ++j) {
      const int z = t / (nt_n * nt_m);
      const uint32_t idesc = idesc0 | ((uint32_t)(tc_n_eff(p, (t % nt_n) * TC_BN, 32) >> 3) << 17);
      const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
      const int buf = j & 1;
      tc_mbar_wait(&bar_acc_empty[buf], ((j >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_main = tmem + (uint32_t)buf * (2 * TC_BN), d_corr = (p.debug & 16) ? d_main : d_main + TC_BN;  // 16: precision experiment, one accumulator
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TC2_STAGES;
        tc_mbar_wait((p.split_a | p.split_b) ? &bar_split[s] : &bar_full[s], (it / TC2_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem0 + s * TC2_STAGE_BYTES;
        const uint32_t sA = st, sAlo = st + TC2_A_BYTES, sB = st + 2 * TC2_A_BYTES, sBlo = sB + TC2_B_BYTES;
        const uint32_t a0 = tc_desc_lo<AMAJ>(sA), al0 = tc_desc_lo<AMAJ>(sAlo), b0 = tc_desc_lo<BMAJ>(sB), bl0 = tc_desc_lo<BMAJ>(sBlo);
        const uint32_t ahi = tc_desc_hi<AMAJ>(), bhi = tc_desc_hi<BMAJ>();
#pragma unroll
        for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
          const uint32_t ka = k8 * tc_desc_k8_step<AMAJ>(), kb8 = k8 * tc_desc_k8_step<BMAJ>();
          const uint32_t first = (kb == kb0 && k8 == 0) ? 0u : 1u;
          tc_mma_lh<2>(d_main, a0 + ka, ahi, b0 + kb8, bhi, idesc, first);
          const uint32_t firstc = (p.debug & 16) ? 1u : first;  // shared accumulator: the correction products always accumulate
          if (p.has_alo) tc_mma_lh<2>(d_corr, al0 + ka, ahi, b0 + kb8, bhi, idesc, firstc);
          if (p.has_blo) tc_mma_lh<2>(d_corr, a0 + ka, ahi, bl0 + kb8, bhi, idesc, (firstc || p.has_alo) ? 1u : 0u);
        }
        tc2_commit(&bar_empty[s]);  // frees the stage in both CTAs once these MMAs have read it
      }
      tc2_commit(&bar_acc_full[buf]);  // accumulator set complete, both CTAs
    }
   }
  }
  } else if (warp >= 12) {
    TC_REG_DEC();
    // ------------------------------------------------------------ operand splitters: this CTA's tiles, 4 warps; 8 arrivals (both CTAs)
    // on the leader's split[s]
    if (p.split_a | p.split_b) {
      uint8_t* const base = tc_smem_raw + (smem0 - tc_smem_u32(tc_smem_raw));
      const int tid = threadIdx.x - 384;
      int it = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        const int z = t / (nt_n * nt_m);
        const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % TC2_STAGES;
          tc_mbar_wait(&bar_full[s], (it / TC2_STAGES) & 1);
          uint8_t* const st = base + (size_t)s * TC2_STAGE_BYTES;
          if (p.split_a && !(p.debug & 4)) tc_split_tile(st, st + TC2_A_BYTES, TC2_A_BYTES / 16, tid);
          if (p.split_b && !(p.debug & 4)) tc_split_tile(st + 2 * TC2_A_BYTES, st + 2 * TC2_A_BYTES + TC2_B_BYTES, TC2_B_BYTES / 16, tid);
          if (!(p.debug & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tc_mapa(tc_smem_u32(&bar_split[s]), 0)) : "memory");
        }
      }
    }
  } else {
    TC_REG_INC();
    // ------------------------------------------------------------ epilogue: this CTA's 128 rows, 8 warps
    float* const stg = reinterpret_cast<float*>(tc_smem_raw + (smem0 - tc_smem_u32(tc_smem_raw)) + TC2_STAGES * TC2_STAGE_BYTES) + (warp - 4) * (TC2_STG_BUFS * 32 * 32);
    int j = 0, stg_turn = 0;
    for (int t = pair; t < ntiles; t += npairs, ++j) {
      const int n0 = (t % nt_n) * TC_BN, m0 = ((t / nt_n) % nt_m) * (2 * TC_BM) + (int)rank * TC_BM, z = t / (nt_n * nt_m);
      const int buf = j & 1;
      if (p.direct == 2) {
        tc_epilogue_tma<TC_BN, TC2_STG_BUFS>(p, &mapC, &mapClo, tmem, buf, warp, lane, m0, n0, z, stg, stg_turn, &bar_acc_full[buf], (j >> 1) & 1,
                                             tc_mapa(tc_smem_u32(&bar_acc_empty[buf]), 0), true);
      } else {
        tc_mbar_wait(&bar_acc_full[buf], (j >> 1) & 1);
        tc_epilogue_tile<TC_BN>(p, tmem, buf, warp, lane, m0, n0, z, stg, tc_mapa(tc_smem_u32(&bar_acc_empty[buf]), 0), true);
      }
    }
    if (p.direct == 2 && lane == 0) tc_bulk_wait_read();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  tc_cluster_sync();  // the peer may still be signalling this CTA's barriers / reading its shared memory through the pair's MMAs
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

// ================================================================== CTA-pair variant with 256-column MMAs (experiment, off by default)
// What bounds k_gemm_tc2 on the 512..752-wide layers is the L2 -> SM operand stream: every k-block (768 tensor cycles for the three
// MMAs of its four k8-steps) a CTA pulls 48 KB through TMA (A, A_lo: 2 x 16 KB; its halves of B, B_lo: 2 x 8 KB) = 64 B/cycle per
// SM, and the chip's L2 slices deliver ~6300 B/cycle in total = 42.6 B/cycle per SM (B300_MICROARCH.md, LTS throughput cap; r1's
// ncu capture: 10.9 TB/s L2 -> SM) - a ceiling of 66 % on the tensor pipe, which is what ncu shows (60-72 %).  TMA multicast does not
// lift it (the L2 de-duplication window only pays from cluster size 8).  An MMA of N = 256 reuses each A tile for twice the
// columns: 64 KB per 1536 tensor cycles = 41.7 B/cycle, at the cap.  Price: main + correction accumulators of 256 columns fill
// TMEM (512 columns), so a tile's epilogue no longer overlaps the next tile's MMAs - and that costs more than the operand
// stream gains (see dtc_gemm_tc3_mode below).
//   * pair tile 256 x 256; each CTA stages its 128 rows of A / A_lo and 128 of the 256 B / B_lo rows: 64 KB per stage, 3 stages;
//   * barriers as in k_gemm_tc2 with a single accumulator set; n-tiles of the 693 / 752-wide layers end in an MMA of N = 192 / 256.
#define TC3_BN 256
#define TC3_STAGES 3
#define TC3_STAGE_BYTES (4 * TC_TILE_BYTES)
#define TC3_SMEM_BYTES (TC3_STAGES * TC3_STAGE_BYTES + TC_STG_BYTES + 1024)

__device__ __forceinline__ int tc3_n_eff(const TcParams& p, int n0) {
  const int live = min(TC3_BN, p.N - n0);
  return min(TC3_BN, (live + 31) / 32 * 32);
}

template <int AMAJ, int BMAJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
k_gemm_tc3(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapB,
           const __grid_constant__ CUtensorMap mapBlo, const __grid_constant__ CUtensorMap mapC,
           const __grid_constant__ CUtensorMap mapClo, const TcParams p) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC3_STAGES], bar_empty[TC3_STAGES], bar_acc_full, bar_acc_empty;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (tc_smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc_cluster_ctarank();
  const int nt_n = (p.N + TC3_BN - 1) / TC3_BN, nt_m = (p.M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int ntiles = nt_n * nt_m * p.splits;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  tc_launch_dependents();

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC3_STAGES; ++s) { tc_mbar_init(&bar_full[s], 1); tc_mbar_init(&bar_empty[s], 1); }
    tc_mbar_init(&bar_acc_full, 1);
    tc_mbar_init(&bar_acc_empty, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  tc_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  tc_grid_dependency_wait();

  if (warp < 4) {
  TC_REG_DEC();
  if (warp == 0) {
   if (tc_elect_one()) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t bytes_cta = (uint32_t)(TC_TILE_BYTES * (2 + p.has_alo + p.has_blo));
    int it = 0;
    for (int t = pair; t < ntiles; t += npairs) {
      const int nt0 = (t % nt_n) * TC3_BN;
      const int n0 = nt0 + (int)rank * (tc3_n_eff(p, nt0) / 2), m0 = ((t / nt_n) % nt_m) * (2 * TC_BM) + (int)rank * TC_BM;
      const int z = t / (nt_n * nt_m);
      const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TC3_STAGES;
        tc_mbar_wait(&bar_empty[s], ((it / TC3_STAGES) & 1) ^ 1);
        const uint32_t full = tc_mapa(tc_smem_u32(&bar_full[s]), 0);  // the leader's barrier
        if (rank == 0) tc_mbar_expect_tx(&bar_full[s], 2 * bytes_cta);
        const uint32_t st = smem0 + s * TC3_STAGE_BYTES;
        const uint32_t sA = st, sAlo = st + TC_TILE_BYTES, sB = st + 2 * TC_TILE_BYTES, sBlo = st + 3 * TC_TILE_BYTES;
        if (AMAJ == 0) {
          tc2_tma_2d(sA, &mapA, full, kb * TC_BK, m0);
          if (p.has_alo) tc2_tma_2d(sAlo, &mapAlo, full, kb * TC_BK, m0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tc2_tma_2d(sA + j * 4096, &mapA, full, m0 + 32 * j, kb * TC_BK);
            if (p.has_alo) tc2_tma_2d(sAlo + j * 4096, &mapAlo, full, m0 + 32 * j, kb * TC_BK);
          }
        }
        if (BMAJ == 0) {
          tc2_tma_2d(sB, &mapB, full, kb * TC_BK, n0);
          if (p.has_blo) tc2_tma_2d(sBlo, &mapBlo, full, kb * TC_BK, n0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tc2_tma_2d(sB + j * 4096, &mapB, full, n0 + 32 * j, kb * TC_BK);
            if (p.has_blo) tc2_tma_2d(sBlo + j * 4096, &mapBlo, full, n0 + 32 * j, kb * TC_BK);
          }
        }
      }
    }
   }
  } else if (warp == 1 && rank == 0) {
   if (tc_elect_one()) {
    // ------------------------------------------------------------ MMA issuer (one thread of the leader CTA), M = 256, N <= 256
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)AMAJ << 15) | ((uint32_t)BMAJ << 16) |
                            ((uint32_t)((2 * TC_BM) >> 4) << 24);
    int it = 0, j = 0;
    for (int t = pair; t < ntiles; t += npairs, ++j) {
      const int z = t / (nt_n * nt_m);
      const uint32_t idesc = idesc0 | ((uint32_t)(tc3_n_eff(p, (t % nt_n) * TC3_BN) >> 3) << 17);
      const int kb0 = z * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
      tc_mbar_wait(&bar_acc_empty, (j & 1) ^ 1);  // both CTAs' epilogues have drained the accumulators of the previous tile
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_main = tmem, d_corr = tmem + TC3_BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % TC3_STAGES;
        tc_mbar_wait(&bar_full[s], (it / TC3_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem0 + s * TC3_STAGE_BYTES;
        const uint32_t a0 = tc_desc_lo<AMAJ>(st), al0 = tc_desc_lo<AMAJ>(st + TC_TILE_BYTES);
        const uint32_t b0 = tc_desc_lo<BMAJ>(st + 2 * TC_TILE_BYTES), bl0 = tc_desc_lo<BMAJ>(st + 3 * TC_TILE_BYTES);
        const uint32_t ahi = tc_desc_hi<AMAJ>(), bhi = tc_desc_hi<BMAJ>();
#pragma unroll
        for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
          const uint32_t ka = k8 * tc_desc_k8_step<AMAJ>(), kb8 = k8 * tc_desc_k8_step<BMAJ>();
          const uint32_t first = (kb == kb0 && k8 == 0) ? 0u : 1u;
          tc_mma_lh<2>(d_main, a0 + ka, ahi, b0 + kb8, bhi, idesc, first);
          if (p.has_alo) tc_mma_lh<2>(d_corr, al0 + ka, ahi, b0 + kb8, bhi, idesc, first);
          if (p.has_blo) tc_mma_lh<2>(d_corr, a0 + ka, ahi, bl0 + kb8, bhi, idesc, (first || p.has_alo) ? 1u : 0u);
        }
        tc2_commit(&bar_empty[s]);
      }
      tc2_commit(&bar_acc_full);
    }
   }
  }
  } else if (warp >= 12) {
    TC_REG_DEC();
  } else {
    TC_REG_INC();
    // ------------------------------------------------------------ epilogue: this CTA's 128 rows x 256 columns, 8 warps
    float* const stg = reinterpret_cast<float*>(tc_smem_raw + (smem0 - tc_smem_u32(tc_smem_raw)) + TC3_STAGES * TC3_STAGE_BYTES) + (warp - 4) * (32 * 32);
    int j = 0, stg_turn = 0;
    for (int t = pair; t < ntiles; t += npairs, ++j) {
      const int n0 = (t % nt_n) * TC3_BN, m0 = ((t / nt_n) % nt_m) * (2 * TC_BM) + (int)rank * TC_BM, z = t / (nt_n * nt_m);
      if (p.direct == 2) {
        tc_epilogue_tma<TC3_BN, 1>(p, &mapC, &mapClo, tmem, 0, warp, lane, m0, n0, z, stg, stg_turn, &bar_acc_full, j & 1,
                                   tc_mapa(tc_smem_u32(&bar_acc_empty), 0), true);
      } else {
        tc_mbar_wait(&bar_acc_full, j & 1);
        tc_epilogue_tile<TC3_BN>(p, tmem, 0, warp, lane, m0, n0, z, stg, tc_mapa(tc_smem_u32(&bar_acc_empty), 0), true);
      }
    }
    if (p.direct == 2 && lane == 0) tc_bulk_wait_read();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  tc_cluster_sync();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

// ------------------------------------------------------------------ host side: tensor-map cache + launcher
typedef CUresult (*PFN_tc_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tc_encode g_encode = nullptr;

struct MapKey {
  const void* p; int rows, k, ld, maj;  // maj: 0 K-major box 32k x 128 rows, 1 MN-major box 32 rows x 32k, 2 K-major box 32k x 64 rows
  bool operator==(const MapKey& o) const { return p == o.p && rows == o.rows && k == o.k && ld == o.ld && maj == o.maj; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.p * 1000003u;
    h ^= ((size_t)k.rows << 1) ^ ((size_t)k.k << 21) ^ ((size_t)k.ld << 41) ^ (size_t)k.maj;
    return h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;  // the cache is shared by every host thread that launches GEMMs

// operand X(r, k): maj 0 -> X = base[r*ld + k] (box 32 k x 128 r); maj 1 -> X = base[k*ld + r] (box 32 r x 32 k)
static int tc_get_map(const float* base, int rows, int k, int ld, int maj, CUtensorMap* out) {
  MapKey key{base, rows, k, ld, maj};
  std::lock_guard<std::mutex> lock(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return DTC_OK; }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DTC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) DTC_FAIL(DTC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = (PFN_tc_encode)fn;
  }
  cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2], es[2] = {1, 1};
  CUtensorMapSwizzle swz;
  if (maj == 0 || maj == 2) { dims[0] = k; dims[1] = rows; box[0] = TC_BK; box[1] = maj == 0 ? 128 : 64; swz = CU_TENSOR_MAP_SWIZZLE_128B; }
  else { dims[0] = rows; dims[1] = k; box[0] = 32; box[1] = TC_BK; swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; }
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DTC_FAIL(DTC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%d x %d] ld %d major %d", (int)r, rows, k, ld, maj);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return DTC_OK;
}

// output of the TMA-store epilogue: box 32 columns x 32 rows, SWIZZLE_128B.  planes == 0: C[rows][ld] with `cols` live columns (the
// unit clips the ragged edges); planes > 0: the split-K workspace [planes][rows][cols] as a 3-D tensor (a box never crosses a plane)
static int tc_get_out_map(const float* base, int rows, int cols, int ld, int planes, CUtensorMap* out) {
  MapKey key{base, rows, cols, planes > 0 ? planes : ld, planes > 0 ? 4 : 3};
  std::lock_guard<std::mutex> lock(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return DTC_OK; }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DTC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) DTC_FAIL(DTC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = (PFN_tc_encode)fn;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(planes > 0 ? planes : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)(planes > 0 ? cols : ld) * sizeof(float), (cuuint64_t)rows * cols * sizeof(float)};
  cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, planes > 0 ? 3 : 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DTC_FAIL(DTC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for output [%d x %d] ld %d planes %d", (int)r, rows, cols, ld, planes);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return DTC_OK;
}

bool dtc_gemm_tc_eligible(const GemmArgs& a) {
  // N down to 32 pays off even though the tile pads it to 128: the 35..64-wide CENet layers run 2-3x faster than on the SIMT path
  if (a.M < 64 || a.N < 32 || a.K < 8) return false;
  if (!(a.A_lo || a.a_split) || !(a.B_lo || a.b_split)) return false;  // fp32-grade results need both companions; otherwise the FP32 SIMT path runs
  if (a.a_kc != true && a.b_kc == true) return false;  // (MN-major A, K-major B) is not used by the learner
  return true;
}

static int g_tc_pair = -1;  // CTA-pair kernel for tile-rich shapes (env DTC_GEMM_PAIR=0 disables)
static int tc_pair_mode() {
  if (g_tc_pair < 0) { const char* e = getenv("DTC_GEMM_PAIR"); g_tc_pair = (e && e[0] == '0') ? 0 : 1; }
  return g_tc_pair;
}
extern "C" void dtc_set_gemm_pair(int on) { g_tc_pair = on ? 1 : 0; }
extern "C" int dtc_get_gemm_pair(void) { return tc_pair_mode(); }

// Programmatic dependent launch of the tensor-core GEMMs: OFF by default (env DTC_PDL=1 enables).  Measured on the training step
// (gpurun_out/r2n): 89.3 / 90.2 ms with it, 88.6 / 88.7 ms without - these persistent CTAs fill an SM, so a dependent grid can only
// start on the SMs the running one leaves to the side streams, where it then spins in griddepcontrol.wait instead of letting the
// side streams' small kernels run.
static int g_tc_pdl = -1;
static int tc_pdl_on() {
  if (g_tc_pdl < 0) { const char* e = getenv("DTC_PDL"); g_tc_pdl = !e ? 0 : e[0] == '1' ? 1 : e[0] == '2' ? 2 : e[0] == '3' ? 3 : 0; }
  return g_tc_pdl;  // 2: only grids that leave SMs free (the 4096-row rollout shapes)
}
template <typename K>
static int tc_launch_pdl(K kernel, dim3 grid, int smem, cudaStream_t st, const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB,
                         const CUtensorMap& mBlo, const CUtensorMap& mC, const CUtensorMap& mClo, const TcParams& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (tc_pdl_on() == 1 || (tc_pdl_on() == 2 && grid.x <= 128) || (tc_pdl_on() == 3 && p.pdl_late)) ? 1 : 0;
  DTC_CUDA(cudaLaunchKernelEx(&cfg, kernel, mA, mAlo, mB, mBlo, mC, mClo, p));
  return DTC_OK;
}

template <int AMAJ, int BMAJ>
static int tc2_launch_t(const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB, const CUtensorMap& mBlo, const CUtensorMap& mC,
                        const CUtensorMap& mClo, const TcParams& p,
                        dim3 grid, cudaStream_t st) {
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device
  int dev_ = 0; cudaGetDevice(&dev_);
  bool& attr_set = attr_set_dev[dev_ & 63];
  if (!attr_set) {
    DTC_CUDA(cudaFuncSetAttribute(k_gemm_tc2<AMAJ, BMAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
    attr_set = true;
  }
  return tc_launch_pdl(k_gemm_tc2<AMAJ, BMAJ>, grid, TC2_SMEM_BYTES, st, mA, mAlo, mB, mBlo, mC, mClo, p);
}

template <int AMAJ, int BMAJ>
static int tc3_launch_t(const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB, const CUtensorMap& mBlo, const CUtensorMap& mC,
                        const CUtensorMap& mClo, const TcParams& p,
                        dim3 grid, cudaStream_t st) {
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device
  int dev_ = 0; cudaGetDevice(&dev_);
  bool& attr_set = attr_set_dev[dev_ & 63];
  if (!attr_set) {
    DTC_CUDA(cudaFuncSetAttribute(k_gemm_tc3<AMAJ, BMAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES));
    attr_set = true;
  }
  return tc_launch_pdl(k_gemm_tc3<AMAJ, BMAJ>, grid, TC3_SMEM_BYTES, st, mA, mAlo, mB, mBlo, mC, mClo, p);
}
// 256-column pair tiles: OFF by default (env DTC_GEMM_TC3=1 enables).  The kernel is correct (bit-identical to the other two:
// tests/test_learner_gpu.py::test_gemm_cta_pair_matches_single runs it when enabled) but slower on every learner shape
// (gpurun_out/r2p, 24 576 rows: 512x512 fwd 152 vs 178 TFLOP/s, dgrad 161 vs 198, 512x693 wgrad 140 vs 168): with TMEM full the
// store burst of each tile (262 KB per CTA, all pairs in lock-step) is exposed instead of hiding under the next tile's MMAs.
static int g_tc3 = -1;
int dtc_gemm_tc3_mode() {
  if (g_tc3 < 0) { const char* e = getenv("DTC_GEMM_TC3"); g_tc3 = (e && e[0] == '1') ? 1 : 0; }
  return g_tc3;
}
// 256-column tiles pay off when the output is at least two thirds of a tile wide and there are enough tiles for every SM pair
bool dtc_gemm_tc3_shape(int M, int N, int splits, int pairs) {
  return dtc_gemm_tc3_mode() && N >= 176 && M > TC_BM && ceil_div(N, TC3_BN) * ceil_div(M, 2 * TC_BM) * splits >= pairs;
}

template <int AMAJ, int BMAJ>
static int tc_launch_t(const CUtensorMap& mA, const CUtensorMap& mAlo, const CUtensorMap& mB, const CUtensorMap& mBlo, const CUtensorMap& mC,
                        const CUtensorMap& mClo, const TcParams& p,
                       dim3 grid, cudaStream_t st) {
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device
  int dev_ = 0; cudaGetDevice(&dev_);
  bool& attr_set = attr_set_dev[dev_ & 63];
  if (!attr_set) {
    DTC_CUDA(cudaFuncSetAttribute(k_gemm_tc<AMAJ, BMAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  return tc_launch_pdl(k_gemm_tc<AMAJ, BMAJ>, grid, TC_SMEM_BYTES, st, mA, mAlo, mB, mBlo, mC, mClo, p);
}

void k_splitk_reduce_launch(const float* ws, float* C, float* C_lo, int M, int N, int ldc, int splits, int accumulate, cudaStream_t st);

int dtc_gemm_tc_launch(GemmArgs a, cudaStream_t st) {
  TcParams p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.nkb = ceil_div(a.K, TC_BK);
  int splits = a.splits < 1 ? 1 : a.splits;
  if (a.ws && a.epi == EPI_STORE && splits < ceil_div(p.nkb, TC_MAX_KB)) splits = ceil_div(p.nkb, TC_MAX_KB);
  if (splits > p.nkb) splits = p.nkb;
  p.kb_per_split = ceil_div(p.nkb, splits);
  splits = ceil_div(p.nkb, p.kb_per_split);
  p.splits = splits;
  if (splits > 1 && (a.epi != EPI_STORE || !a.ws)) DTC_FAIL(DTC_ERR_ARG, "gemm_tc: split-K needs EPI_STORE and a workspace");
  p.C = a.C; p.C_lo = a.C_lo; p.ldc = a.ldc;
  p.bias = a.bias; p.act_src = a.act_src; p.ld_act = a.ld_act; p.epi = a.epi; p.accumulate = a.accumulate ? 1 : 0;
  p.ws = a.ws;
  p.colsum_part = (splits == 1) ? a.colsum_part : nullptr;
  p.colsum_out = (splits == 1) ? a.colsum_out : nullptr; p.colsum_n = a.colsum_n;
  p.split_a = a.a_split ? 1 : 0; p.split_b = a.b_split ? 1 : 0;
  p.has_alo = (a.A_lo || a.a_split) ? 1 : 0; p.has_blo = (a.B_lo || a.b_split) ? 1 : 0;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("DTC_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  // epilogue variant: "staged" (default) or "direct".  Measured (gpurun_out/r2o): direct wins on dgrad / narrow shapes alone (256x512
  // dgrad 124 -> 148 TFLOP/s, 128x256 60 -> 77) but loses on the forward shapes (512x693 184 -> 169) and costs 6 ms per training
  // iteration (88.5 -> 94.5 ms): its half-sector writes load the L2 write path that the step already saturates.
  // DTC_TC_EPI = tma (default: registers -> one shared-memory chunk -> cp.async.bulk.tensor store, see tc_epilogue_tma) | staged | direct
  { static int direct = -1; if (direct < 0) { const char* e = getenv("DTC_TC_EPI"); direct = !e ? 2 : e[0] == 'd' ? 1 : e[0] == 's' ? 0 : 2; } p.direct = direct; }
  // Measured in the training step (gpurun_out/r2q, ms per iteration / in-step TFLOP/s of the pair kernel): tma 86.7 / 138.7, staged
  // 95.8 / 117.1, tma for forward + staged for the mask / fan-in epilogues 88.4 / 127.7.
  { static int lod = -1; if (lod < 0) { const char* e = getenv("DTC_TC_LO"); lod = (e && e[0] == 'd') ? 1 : 0; } p.lo_direct = lod; }
  { static int neff = -1; if (neff < 0) { const char* e = getenv("DTC_TC_NEFF"); neff = e ? atoi(e) : 1; } p.neff = neff; }  // DTC_TC_NEFF=0: always 128-column MMAs
  const int amaj = a.a_kc ? 0 : 1, bmaj = a.b_kc ? 0 : 1;
  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
  // CTA pairs pay off once there are enough 256-row tiles to keep every pair busy
  const int pair_tiles = ceil_div(a.N, TC_BN) * ceil_div(a.M, 2 * TC_BM) * splits;
  static int pair_min = -1;  // fewest pair tiles worth a cluster launch (env DTC_GEMM_PAIR_MIN, default: one per SM pair)
  if (pair_min < 0) { const char* e = getenv("DTC_GEMM_PAIR_MIN"); pair_min = e ? atoi(e) : num_sms / 2; }
  const bool use_pair = tc_pair_mode() && a.M > TC_BM && pair_tiles >= pair_min;
  p.pdl_late = (tc_pdl_on() == 3 && !use_pair && a.M <= 4096 && splits == 1) ? 1 : 0;  // the rollout's GEMM -> GEMM chains
  const bool use_tc3 = use_pair && !a.a_split && !a.b_split && dtc_gemm_tc3_shape(a.M, a.N, splits, pair_min);
  const int bmap = (use_pair && !use_tc3 && bmaj == 0) ? 2 : bmaj;  // tc2 stages 64 B rows per CTA, tc3 128
  CUtensorMap mA, mAlo, mB, mBlo;
  RETURN_IF_ERR(tc_get_map(a.A, a.M, a.K, a.lda, amaj, &mA));
  RETURN_IF_ERR(tc_get_map(a.B, a.N, a.K, a.ldb, bmap, &mB));
  if (a.A_lo && !a.a_split) RETURN_IF_ERR(tc_get_map(a.A_lo, a.M, a.K, a.lda, amaj, &mAlo)); else mAlo = mA;
  if (a.B_lo && !a.b_split) RETURN_IF_ERR(tc_get_map(a.B_lo, a.N, a.K, a.ldb, bmap, &mBlo)); else mBlo = mB;
  // split-K through L2 reductions (default; env DTC_TC_SPLITK=ws restores workspace + reduce kernel): needs the TMA epilogue, no
  // companion output, a destination whose padding columns may be written (see below).  The summation order of the partial tiles is
  // then whatever order the CTAs finish in: fp32-round-off differences from run to run.
  { static int sk = -1; if (sk < 0) { const char* e = getenv("DTC_TC_SPLITK"); sk = (e && e[0] == 'w') ? 0 : 1; }
    p.splitk_atomic = (sk && splits > 1 && p.direct == 2 && !a.C_lo && (!(a.N & 3) || a.ldc == ((a.N + 3) & ~3))) ? 1 : 0; }
  if (p.splitk_atomic && !a.accumulate && !a.c_zeroed) DTC_CUDA(cudaMemsetAsync(a.C, 0, (size_t)(a.M - 1) * a.ldc * sizeof(float) + (size_t)((a.N + 3) & ~3) * sizeof(float), st));
  CUtensorMap mC = mA, mClo = mA;
  // the TMA unit clips a store at 16-byte granularity: with N % 4 != 0 it writes zeros into the columns up to round4(N).  Harmless when
  // those are this matrix's own padding (ldc == round4(N)); a narrower view into a wider buffer keeps the element-exact staged epilogue
  if (p.direct == 2 && (a.N & 3) && splits == 1 && a.ldc != ((a.N + 3) & ~3)) p.direct = 0;
  if (p.direct == 2) {
    const int n4 = (a.N + 3) & ~3;
    if (splits > 1 && !p.splitk_atomic) RETURN_IF_ERR(tc_get_out_map(a.ws, a.M, n4, n4, splits, &mC));
    else {
      RETURN_IF_ERR(tc_get_out_map(a.C, a.M, a.N, a.ldc, 0, &mC));
      if (a.C_lo) RETURN_IF_ERR(tc_get_out_map(a.C_lo, a.M, a.N, a.ldc, 0, &mClo));
    }
  }
  const int ntiles = use_tc3 ? ceil_div(a.N, TC3_BN) * ceil_div(a.M, 2 * TC_BM) * splits
                             : use_pair ? pair_tiles : ceil_div(a.N, TC_BN) * ceil_div(a.M, TC_BM) * splits;
  // persistent grid: the tile list takes R = ceil(tiles / units) rounds whatever happens, so launch only ceil(tiles / R) units -
  // same makespan, and the SMs left over run the side streams' small kernels (nothing can co-reside with these CTAs)
  const int units = use_pair ? num_sms / 2 : num_sms;
  const int rounds = ceil_div(ntiles, units);
  const int used = ceil_div(ntiles, rounds);
  dim3 grid(use_pair ? 2 * used : used);
  dtc_prof_begin(st, use_pair ? 2 : 0, 2.0 * a.M * a.N * a.K);
  dtc_prof_tag(a.M, a.N, a.K, (use_tc3 ? 20 : use_pair ? 10 : 0) + amaj * 2 + bmaj + 1000 * splits);
  int rc;
  if (use_tc3) {
    if (amaj == 0 && bmaj == 0) rc = tc3_launch_t<0, 0>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
    else if (amaj == 0 && bmaj == 1) rc = tc3_launch_t<0, 1>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
    else if (amaj == 1 && bmaj == 1) rc = tc3_launch_t<1, 1>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
    else DTC_FAIL(DTC_ERR_ARG, "gemm_tc: unsupported operand layout");
  } else if (use_pair) {
    if (amaj == 0 && bmaj == 0) rc = tc2_launch_t<0, 0>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
    else if (amaj == 0 && bmaj == 1) rc = tc2_launch_t<0, 1>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
    else if (amaj == 1 && bmaj == 1) rc = tc2_launch_t<1, 1>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
    else DTC_FAIL(DTC_ERR_ARG, "gemm_tc: unsupported operand layout");
  } else if (amaj == 0 && bmaj == 0) rc = tc_launch_t<0, 0>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
  else if (amaj == 0 && bmaj == 1) rc = tc_launch_t<0, 1>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
  else if (amaj == 1 && bmaj == 1) rc = tc_launch_t<1, 1>(mA, mAlo, mB, mBlo, mC, mClo, p, grid, st);
  else DTC_FAIL(DTC_ERR_ARG, "gemm_tc: unsupported operand layout");
  if (rc) return rc;
  DTC_CHECK_LAUNCH("k_gemm_tc");
  if (splits > 1 && !p.splitk_atomic) k_splitk_reduce_launch(a.ws, a.C, a.C_lo, a.M, a.N, a.ldc, splits, a.accumulate ? 1 : 0, st);
  dtc_prof_end(st);
  return DTC_OK;
}
