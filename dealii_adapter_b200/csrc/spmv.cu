// K4: FP64 block-row SpMV y = A x with an optional fused dot product (x . A x) — the `vmult`
// inside SolverCG (nonlinear_elasticity.cc:1184, linear_elasticity.cc:551) and the explicit
// vmults of assemble_rhs (linear_elasticity.cc:411-419).
//
// Storage (gf_context.h): per node row A with nb blocks, `dim` scalar rows of nb*dim doubles
// (padded to an even count), all sharing ONE column-index stream: 8 B per value + 4/dim^2 B of
// index per value instead of CSR's 8+4. HBM-bound; algorithmic bytes per launch =
// 8*n_val + 4*n_blocks + 4*(n_rows+1) + 8*(n_rows+1) + 8*n_local(x) + 8*n_owned(y).
//
// Kernels (GF_OPT_SPMV_KERNEL; all return bitwise identical y: same tiles, same per-row
// summation order). TILES are runs of consecutive rows (<= 6144 values + their column indices +
// per-row records), built once in pattern.cu; every TMA kernel is persistent (one CTA per SM) and
// warp-specialised, operands arrive by TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx,
// SASS UBLKCP) in mbarrier-guarded shared-memory rings, so the bytes in flight live in shared
// memory, not registers (the LDG version was latency-bound at 32 warps/SM,
// profiles/r01_spmv_ncu_summary.md):
//   5  spmv_tma_kernel: one 3-stage ring carries a tile through TMA -> x gather (8 warps) ->
//      FMA (8 warps). 0.624 ms on the cfg3 tangent (5.33 TB/s).
//   3  spmv_tma2_kernel<.., GW, GG, CW>: column indices / row records / gathered x in their own
//      deeper ring running ahead of the value ring; GG groups of gather warps work on consecutive
//      tiles; CW consumer warps. 8+16 warps (kind 3): 0.549 ms = 6.05 TB/s = 92 % of the measured
//      copy peak - the consumer warps, not HBM or the gather, bounded the 8-consumer kernels
//      (profiles/r01_spmv_variants_ncu.md).
//   6  kind 3 with the three (two) scalar-row sums of a block row reduced by ONE transposed
//      butterfly (kernel_utils.cuh::warp_sum_rows; bitwise the same sums): 0.517 ms = 6.43 TB/s =
//      98 % of the measured copy peak (profiles/r02_spmv_kernel_kinds.jsonl).
//   0  (default) kind 6 for every launch, with and without the fused dot product.
//   1  spmv_kernel (LDG): one warp per row with streaming loads; also the fallback when a row
//      does not fit a tile.
// (The kernels still carry a DOT template parameter: a fused x . A x with one partial per CTA. The
// CG stopped using it in round 2 - partial sums per CTA depend on the partition - and DOT = true
// is no longer instantiated.)
//
// Value type VT: double for every operator application whose result the reference defines (CG
// vmult, assemble_rhs vmults). VT = float streams a single-precision COPY of the same array
// (same indexing, 4 B per value) and is used ONLY inside the multigrid V-cycle when
// GF_OPT_MG_MATRIX_PRECISION = 1: the preconditioner replaces the reference's SSOR and may be
// any fixed SPD operator; vectors, accumulation and the outer CG stay FP64
// (GF_OPT_MG_MATRIX_PRECISION = 2, experimental: x staging and accumulation in FP32 as well).
#include "gf_context.h"
#include "kernel_utils.cuh"

namespace gf
{
  namespace
  {
    constexpr int SPMV_THREADS = 256;

    __device__ __forceinline__ double2 ld_stream2(const double *p)
    {
      return __ldcs(reinterpret_cast<const double2 *>(p));
    }
    __device__ __forceinline__ double2 ld_stream2(const float *p)
    {
      const float2 v = __ldcs(reinterpret_cast<const float2 *>(p));
      return make_double2(double(v.x), double(v.y));
    }
    // two consecutive values from shared memory, widened to FP64
    __device__ __forceinline__ double2 lds2(const double *p)
    {
      return *reinterpret_cast<const double2 *>(p);
    }
    __device__ __forceinline__ double2 lds2(const float *p)
    {
      const float2 v = *reinterpret_cast<const float2 *>(p);
      return make_double2(double(v.x), double(v.y));
    }

    template <int DIM, bool DOT, typename VT>
    __global__ void __launch_bounds__(SPMV_THREADS)
      spmv_kernel(const int64_t n_rows, const int32_t *__restrict__ brow_ptr,
                  const int64_t *__restrict__ val_ptr, const int32_t *__restrict__ bcol,
                  const VT *__restrict__ val, const double *__restrict__ x,
                  double *__restrict__ y, double *__restrict__ partials, const int *status)
    {
      if (status != nullptr && *status != 0)
        return;
      __shared__ double red[32];
      const int         lane = threadIdx.x & 31;
      const int64_t     warp_global = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      const int64_t     n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
      double            dot = 0.0;
      for (int64_t A = warp_global; A < n_rows; A += n_warps)
        {
          const int32_t b0     = brow_ptr[A];
          const int     ne     = (brow_ptr[A + 1] - b0) * DIM;
          const int64_t vbase  = val_ptr[A];
          const int     stride = int((val_ptr[A + 1] - vbase) / DIM);
          const VT *    vrow   = val + vbase;
          const int32_t *crow  = bcol + b0;
          double        acc[DIM];
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            acc[r] = 0.0;
#pragma unroll 2
          for (int e = 2 * lane; e < ne; e += 64)
            {
              const int    e1 = e + 1;
              const int    k0 = e / DIM, k1 = e1 / DIM;
              const double x0 = __ldg(x + int64_t(__ldcs(crow + k0)) * DIM + (e - k0 * DIM));
              double       x1 = 0.0;
              if (e1 < ne)
                x1 = __ldg(x + int64_t(__ldcs(crow + k1)) * DIM + (e1 - k1 * DIM));
#pragma unroll
              for (int r = 0; r < DIM; ++r)
                {
                  const double2 v = ld_stream2(vrow + r * stride + e);
                  acc[r]          = fma(v.x, x0, acc[r]);
                  acc[r]          = fma(v.y, x1, acc[r]);
                }
            }
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            acc[r] = warp_sum(acc[r]);
          if (lane < DIM)
            {
              double yr = acc[0];
#pragma unroll
              for (int r = 1; r < DIM; ++r)
                if (lane == r)
                  yr = acc[r];
              y[A * DIM + lane] = yr;
              if (DOT)
                dot = fma(yr, x[A * DIM + lane], dot);
            }
        }
      if (DOT)
        {
          dot = warp_sum(dot);
          if (lane == 0)
            red[threadIdx.x >> 5] = dot;
          __syncthreads();
          if (threadIdx.x < 32)
            {
              double v = threadIdx.x < (SPMV_THREADS >> 5) ? red[threadIdx.x] : 0.0;
              v        = warp_sum(v);
              if (threadIdx.x == 0)
                partials[blockIdx.x] = v;
            }
        }
    }


#ifndef GF_CUDA_EMULATION // TMA / mbarrier kernels: hardware only (tests/cuda_emu runs the LDG kernel)
    // ------------------------------------------------------------------------------------------
    // TMA-tiled kernel: producer warp (TMA) -> gather warps (x) -> consumer warps (FMA)
    // ------------------------------------------------------------------------------------------
    template <int DIM, typename VT>
    struct TmaCfg
    {
      static constexpr int TILE_V       = DIM == 3 ? 6144 : 4096; // values per tile
      static constexpr int TILE_C       = ((TILE_V / (DIM * DIM)) + 8 + 3) & ~3; // column indices
      // FP32 stages are half the size: one more stage keeps the same bytes in flight per SM
      static constexpr int STAGES       = sizeof(VT) == 4 ? 4 : 3;
      static constexpr int GATHER_WARPS = 8;
      static constexpr int CONS_WARPS   = 8;
      static constexpr int THREADS      = (1 + GATHER_WARPS + CONS_WARPS) * 32;
      // FP32: a tile starts at an even element, i.e. 8-byte but not always 16-byte aligned; the
      // bulk copy then starts up to 2 elements early (skew) and is rounded up to 16 bytes
      static constexpr int VAL_BYTES    = (TILE_V * int(sizeof(VT)) + (sizeof(VT) == 4 ? 16 : 0) + 127) & ~127;
      static constexpr int COL_BYTES    = (TILE_C * 4 + 127) & ~127;
      static constexpr int META_BYTES   = SPMV_META * 8; // bytes copied
      static constexpr int META_PAD     = (META_BYTES + 127) & ~127;
      static constexpr int XG_BYTES     = (TILE_C * DIM * 8 + 127) & ~127;
      static constexpr int STAGE_BYTES  = VAL_BYTES + COL_BYTES + META_PAD + XG_BYTES;
      static constexpr int SMEM_BYTES   = STAGES * STAGE_BYTES + 3 * STAGES * 8 + 64;
      static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    };

    // TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
    __device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes,
                                             uint32_t bar)
    {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], "
                   "%2, [%3];" ::"r"(dst),
                   "l"(src), "r"(bytes), "r"(bar)
                   : "memory");
    }

    template <int DIM, bool DOT, typename VT>
    __global__ void __launch_bounds__(TmaCfg<DIM, VT>::THREADS, 1)
      spmv_tma_kernel(const int n_tiles, const TileDesc *__restrict__ tile_desc,
                      const uint2 *__restrict__ tile_meta, const int32_t *__restrict__ bcol,
                      const VT *__restrict__ val, const double *__restrict__ x,
                      double *__restrict__ y, double *__restrict__ partials, const int *status)
    {
      using C = TmaCfg<DIM, VT>;
      if (status != nullptr && *status != 0)
        return;
      extern __shared__ __align__(128) unsigned char smem[];
      uint64_t *bars  = reinterpret_cast<uint64_t *>(smem + C::STAGES * C::STAGE_BYTES);
      uint64_t *full  = bars;                 // TMA bytes landed          (count 1 + tx)
      uint64_t *xfull = bars + C::STAGES;     // x gathered               (count GATHER_WARPS)
      uint64_t *empty = bars + 2 * C::STAGES; // consumers done with tile (count CONS_WARPS)
      __shared__ double red[C::CONS_WARPS];
      const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
      if (tid == 0)
        {
          for (int s = 0; s < C::STAGES; ++s)
            {
              mbar_init(smem_u32(&full[s]), 1);
              mbar_init(smem_u32(&xfull[s]), C::GATHER_WARPS);
              mbar_init(smem_u32(&empty[s]), C::CONS_WARPS);
            }
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
      __syncthreads();
      const int n_my =
        int(blockIdx.x) < n_tiles ? (n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
      double dot = 0.0;

      if (warp == 0)
        {
          // ------------------------------ producer warp: TMA ---------------------------------
          TileDesc d{};
          for (int k = 0; k < n_my; ++k)
            {
              if ((k & 31) == 0)
                {
                  const int kk = k + lane;
                  if (kk < n_my)
                    d = tile_desc[int64_t(blockIdx.x) + int64_t(kk) * gridDim.x];
                }
              const int       src_lane  = k & 31;
              const long long val_off   = __shfl_sync(0xffffffffu, d.val_off, src_lane);
              const int       val_count = __shfl_sync(0xffffffffu, d.val_count, src_lane);
              const int       col_off   = __shfl_sync(0xffffffffu, d.col_off, src_lane);
              const int       col_count = __shfl_sync(0xffffffffu, d.col_count, src_lane);
              const int       s         = k % C::STAGES;
              if (k >= C::STAGES) // stage s was last used by tile k - STAGES
                mbar_wait(smem_u32(&empty[s]), ((k / C::STAGES) - 1) & 1);
              if (lane == 0)
                {
                  const int64_t  t     = int64_t(blockIdx.x) + int64_t(k) * gridDim.x;
                  unsigned char *stage = smem + s * C::STAGE_BYTES;
                  const uint32_t fb    = smem_u32(&full[s]);
                  // 16-byte aligned source and size: FP64 tiles are (even offset and count),
                  // FP32 tiles start `skew` elements early (consumers add the same skew)
                  const int      skew = sizeof(VT) == 4 ? int(val_off & 3) : 0;
                  const uint32_t vb =
                    (uint32_t(skew + val_count) * uint32_t(sizeof(VT)) + 15u) & ~15u;
                  const uint32_t cb = uint32_t(col_count) * 4u;
                  mbar_arrive_expect_tx(fb, vb + cb + uint32_t(C::META_BYTES));
                  bulk_g2s(smem_u32(stage), val + (val_off - skew), vb, fb);
                  bulk_g2s(smem_u32(stage + C::VAL_BYTES), bcol + col_off, cb, fb);
                  bulk_g2s(smem_u32(stage + C::VAL_BYTES + C::COL_BYTES), tile_meta + t * SPMV_META,
                           uint32_t(C::META_BYTES), fb);
                }
              __syncwarp();
            }
        }
      else if (warp <= C::GATHER_WARPS)
        {
          // ---------------- gather warps: x[col] -> shared, many independent loads per lane ----
          const int g = (warp - 1) * 32 + lane;
          for (int k = 0; k < n_my; ++k)
            {
              const int s = k % C::STAGES;
              mbar_wait(smem_u32(&full[s]), (k / C::STAGES) & 1);
              unsigned char *stage = smem + s * C::STAGE_BYTES;
              const int32_t *scol  = reinterpret_cast<const int32_t *>(stage + C::VAL_BYTES);
              const uint2 *  smeta =
                reinterpret_cast<const uint2 *>(stage + C::VAL_BYTES + C::COL_BYTES);
              double *sx =
                reinterpret_cast<double *>(stage + C::VAL_BYTES + C::COL_BYTES + C::META_PAD);
              const int nblk = int(smeta[SPMV_TILE_ROWS + 1].x);
#pragma unroll 4
              for (int j = g; j < nblk; j += C::GATHER_WARPS * 32)
                {
                  const int64_t col = scol[j];
#pragma unroll
                  for (int d0 = 0; d0 < DIM; ++d0)
                    sx[j * DIM + d0] = __ldg(x + col * DIM + d0);
                }
              __syncwarp();
              if (lane == 0)
                mbar_arrive(smem_u32(&xfull[s]));
            }
        }
      else
        {
          // ------------------------------ consumer warps: shared-memory FMA --------------------
          const int cw = warp - 1 - C::GATHER_WARPS;
          for (int k = 0; k < n_my; ++k)
            {
              const int s = k % C::STAGES;
              mbar_wait(smem_u32(&full[s]), (k / C::STAGES) & 1);
              mbar_wait(smem_u32(&xfull[s]), (k / C::STAGES) & 1);
              const unsigned char *stage = smem + s * C::STAGE_BYTES;
              const uint2 *        smeta =
                reinterpret_cast<const uint2 *>(stage + C::VAL_BYTES + C::COL_BYTES);
              // FP32: the copy started smeta[..+1].y = (tile val_off & 3) elements early
              const VT *sval = reinterpret_cast<const VT *>(stage) +
                               (sizeof(VT) == 4 ? int(smeta[SPMV_TILE_ROWS + 1].y) : 0);
              const double *sx = reinterpret_cast<const double *>(stage + C::VAL_BYTES +
                                                                  C::COL_BYTES + C::META_PAD);
              const uint2 hdr    = smeta[SPMV_TILE_ROWS];
              const int   row0   = int(hdr.x);
              const int   n_rows = int(hdr.y);
              for (int row = cw; row < n_rows; row += C::CONS_WARPS)
                {
                  const uint2   m      = smeta[row];
                  const int     ne     = int(m.y >> 16) * DIM;
                  const int     stride = (ne + 1) & ~1;
                  const VT *    v      = sval + m.x;
                  const double *xr     = sx + int(m.y & 0xffffu) * DIM; // x of this row, contiguous
                  const int64_t i      = int64_t(row0 + row) * DIM + lane;
                  double        xi     = 0.0;
                  if (DOT && lane < DIM)
                    xi = __ldg(x + i);
                  double acc[DIM];
#pragma unroll
                  for (int r = 0; r < DIM; ++r)
                    acc[r] = 0.0;
#pragma unroll 3
                  for (int e = 2 * lane; e < ne; e += 64)
                    {
                      const double x0 = xr[e];
                      const double x1 = (e + 1 < ne) ? xr[e + 1] : 0.0;
#pragma unroll
                      for (int r = 0; r < DIM; ++r)
                        {
                          const double2 vv = lds2(v + r * stride + e);
                          acc[r]           = fma(vv.x, x0, acc[r]);
                          acc[r]           = fma(vv.y, x1, acc[r]);
                        }
                    }
#pragma unroll
                  for (int r = 0; r < DIM; ++r)
                    acc[r] = warp_sum(acc[r]);
                  if (lane < DIM)
                    {
                      double yr = acc[0];
#pragma unroll
                      for (int r = 1; r < DIM; ++r)
                        if (lane == r)
                          yr = acc[r];
                      y[i] = yr;
                      if (DOT)
                        dot = fma(yr, xi, dot);
                    }
                }
              __syncwarp();
              if (lane == 0)
                mbar_arrive(smem_u32(&empty[s])); // release the stage
            }
        }
      if (DOT)
        {
          const int cw = warp - 1 - C::GATHER_WARPS;
          dot          = warp_sum(dot);
          if (lane == 0 && cw >= 0)
            red[cw] = dot;
          __syncthreads();
          if (tid < 32)
            {
              double v = tid < C::CONS_WARPS ? red[tid] : 0.0;
              v        = warp_sum(v);
              if (tid == 0)
                partials[blockIdx.x] = v;
            }
        }
    }

    // ------------------------------------------------------------------------------------------
    // Two-ring variant (GF_OPT_SPMV_KERNEL = 2). In the kernel above one stage carries a tile
    // through TMA -> x gather -> FMA, so a stage is busy for the SUM of the three latencies and
    // the bytes in flight per SM are bounded by STAGES / (that sum): the FP32 copy of the same
    // matrix ran in 0.60 ms against 0.62 ms for FP64 (profiles/r01_f32_vcycle_probe_single_ring
    // .json) - latency, not bytes. Here the small per-tile data (column indices, row records,
    // gathered x: ~20 KB) live in their own, deeper ring that runs LEAD tiles ahead of the value
    // ring: the x gather of tile k overlaps the TMA flight of tile k's values, and a value stage
    // is busy only for TMA + FMA.
    //   producer order per step i: cols(i), then values(i - LEAD). cols(i) needs tile i - XSTAGES
    //   consumed, values(j) tile j - VSTAGES; with LEAD = XSTAGES - VSTAGES both waits are for the
    //   same tile, so neither stream throttles the other and the in-order producer cannot deadlock
    //   (everything a waited-for tile needs was issued in an earlier step).
    // ------------------------------------------------------------------------------------------
    template <int DIM, typename VT, int GW, int GG, int CW>
    struct Tma2Cfg
    {
      using B = TmaCfg<DIM, VT>;
      static constexpr int GATHER_WARPS = GW; // gather warps in total
      static constexpr int CONS_WARPS   = CW; // consumer warps
      static constexpr int THREADS      = (1 + GW + CW) * 32;
      static constexpr int VSTAGES      = sizeof(VT) == 4 ? 4 : 3;
      static constexpr int XSTAGES      = sizeof(VT) == 4 ? 6 : 4;
      static constexpr int LEAD         = XSTAGES - VSTAGES;
      // gather groups = tiles gathered concurrently. Must divide XSTAGES: a group then meets the
      // phases of "its" stages consecutively (a skipped phase would alias in the parity wait)
      static constexpr int GROUPS       = GG;
      static_assert(XSTAGES % GROUPS == 0 && GW % GROUPS == 0, "groups");
      static_assert(CW <= 32, "the dot-product tail reduces the consumer warps in one warp");
      static constexpr int XSTAGE_BYTES = B::COL_BYTES + B::META_PAD + B::XG_BYTES;
      static constexpr int RING_BYTES   = VSTAGES * B::VAL_BYTES + XSTAGES * XSTAGE_BYTES;
      static constexpr int N_BARS       = 2 * VSTAGES + 3 * XSTAGES;
      static constexpr int SMEM_BYTES   = RING_BYTES + N_BARS * 8 + 64;
      static_assert(SMEM_BYTES + 64 <= 227 * 1024, "shared memory budget");
    };

    // XT = type of the staged x values and of the accumulators: double everywhere except the
    // all-single-precision operator of the V-cycle (GF_OPT_MG_MATRIX_PRECISION = 2: VT = XT = float,
    // FFMA with two accumulators per scalar row; y is written back as double)
    // TR: the three (two) scalar-row sums of a block row by ONE transposed butterfly
    // (kernel_utils.cuh::warp_sum_rows, bitwise the same sums in 6 instead of 15 shuffle rounds)
    template <int DIM, bool DOT, typename VT, typename XT, int GW, int GG, int CW, bool TR = false>
    __global__ void __launch_bounds__((1 + GW + CW) * 32, 1)
      spmv_tma2_kernel(const int n_tiles, const TileDesc *__restrict__ tile_desc,
                       const uint2 *__restrict__ tile_meta, const int32_t *__restrict__ bcol,
                       const VT *__restrict__ val, const double *__restrict__ x,
                       double *__restrict__ y, double *__restrict__ partials, const int *status)
    {
      using B = TmaCfg<DIM, VT>;
      using C = Tma2Cfg<DIM, VT, GW, GG, CW>;
      static_assert(sizeof(XT) == 8 || (sizeof(VT) == 4 && !DOT),
                    "single-precision x staging only with the FP32 copy and without the fused dot");
      if (status != nullptr && *status != 0)
        return;
      extern __shared__ __align__(128) unsigned char smem[];
      unsigned char *vring  = smem;
      unsigned char *xring  = smem + C::VSTAGES * B::VAL_BYTES;
      uint64_t *     bars   = reinterpret_cast<uint64_t *>(smem + C::RING_BYTES);
      uint64_t *     vfull  = bars;                               // values landed   (1 + tx)
      uint64_t *     vempty = bars + C::VSTAGES;                  // consumers done  (CONS_WARPS)
      uint64_t *     cfull  = bars + 2 * C::VSTAGES;              // cols + records  (1 + tx)
      uint64_t *     xfull  = bars + 2 * C::VSTAGES + C::XSTAGES; // x gathered (warps of a group)
      uint64_t *     xempty = bars + 2 * C::VSTAGES + 2 * C::XSTAGES; // consumers done
      __shared__ double red[C::CONS_WARPS];
      const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
      if (tid == 0)
        {
          for (int s = 0; s < C::VSTAGES; ++s)
            {
              mbar_init(smem_u32(&vfull[s]), 1);
              mbar_init(smem_u32(&vempty[s]), C::CONS_WARPS);
            }
          for (int s = 0; s < C::XSTAGES; ++s)
            {
              mbar_init(smem_u32(&cfull[s]), 1);
              mbar_init(smem_u32(&xfull[s]), C::GATHER_WARPS / C::GROUPS);
              mbar_init(smem_u32(&xempty[s]), C::CONS_WARPS);
            }
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
      __syncthreads();
      const int n_my =
        int(blockIdx.x) < n_tiles ? (n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
      double dot = 0.0;

      if (warp == 0)
        {
          // ------------------------------ producer warp: TMA ---------------------------------
          TileDesc dc{}, dv{}; // lane l caches the descriptor of tile (batch start + l)
          for (int i = 0; i < n_my + C::LEAD; ++i)
            {
              // ---- column indices + row records of tile i ----
              if (i < n_my)
                {
                  if ((i & 31) == 0)
                    {
                      const int kk = i + lane;
                      if (kk < n_my)
                        dc = tile_desc[int64_t(blockIdx.x) + int64_t(kk) * gridDim.x];
                    }
                  const int col_off   = __shfl_sync(0xffffffffu, dc.col_off, i & 31);
                  const int col_count = __shfl_sync(0xffffffffu, dc.col_count, i & 31);
                  const int sx        = i % C::XSTAGES;
                  if (i >= C::XSTAGES)
                    mbar_wait(smem_u32(&xempty[sx]), ((i / C::XSTAGES) - 1) & 1);
                  if (lane == 0)
                    {
                      const int64_t  t     = int64_t(blockIdx.x) + int64_t(i) * gridDim.x;
                      unsigned char *stage = xring + sx * C::XSTAGE_BYTES;
                      const uint32_t fb    = smem_u32(&cfull[sx]);
                      const uint32_t cb    = uint32_t(col_count) * 4u;
                      mbar_arrive_expect_tx(fb, cb + uint32_t(B::META_BYTES));
                      bulk_g2s(smem_u32(stage), bcol + col_off, cb, fb);
                      bulk_g2s(smem_u32(stage + B::COL_BYTES), tile_meta + t * SPMV_META,
                               uint32_t(B::META_BYTES), fb);
                    }
                  __syncwarp();
                }
              // ---- values of tile j = i - LEAD ----
              const int j = i - C::LEAD;
              if (j >= 0)
                {
                  if ((j & 31) == 0)
                    {
                      const int kk = j + lane;
                      if (kk < n_my)
                        dv = tile_desc[int64_t(blockIdx.x) + int64_t(kk) * gridDim.x];
                    }
                  const long long val_off   = __shfl_sync(0xffffffffu, dv.val_off, j & 31);
                  const int       val_count = __shfl_sync(0xffffffffu, dv.val_count, j & 31);
                  const int       sv        = j % C::VSTAGES;
                  if (j >= C::VSTAGES)
                    mbar_wait(smem_u32(&vempty[sv]), ((j / C::VSTAGES) - 1) & 1);
                  if (lane == 0)
                    {
                      const uint32_t fb   = smem_u32(&vfull[sv]);
                      const int      skew = sizeof(VT) == 4 ? int(val_off & 3) : 0;
                      const uint32_t vb =
                        (uint32_t(skew + val_count) * uint32_t(sizeof(VT)) + 15u) & ~15u;
                      mbar_arrive_expect_tx(fb, vb);
                      bulk_g2s(smem_u32(vring + sv * B::VAL_BYTES), val + (val_off - skew), vb, fb);
                    }
                  __syncwarp();
                }
            }
        }
      else if (warp <= C::GATHER_WARPS)
        {
          // ---------------- gather warps: x[col] -> shared --------------------------------------
          // The gather of one tile is ONE round trip to L2 (~1 us under the streaming load), so
          // tiles must overlap: GROUPS groups of warps work on consecutive tiles concurrently, and
          // a lane issues ALL loads of its blocks (UB blocks, fully unrolled) before the first
          // store, so a tile costs a group exactly one round trip.
          constexpr int GL = C::GATHER_WARPS * 32 / C::GROUPS; // lanes per group
          constexpr int UB = (B::TILE_C + GL - 1) / GL;        // blocks per lane and tile
          const int     grp = (warp - 1) / (C::GATHER_WARPS / C::GROUPS);
          const int     g   = ((warp - 1) % (C::GATHER_WARPS / C::GROUPS)) * 32 + lane;
          for (int k = grp; k < n_my; k += C::GROUPS)
            {
              const int sx = k % C::XSTAGES;
              mbar_wait(smem_u32(&cfull[sx]), (k / C::XSTAGES) & 1);
              unsigned char *stage = xring + sx * C::XSTAGE_BYTES;
              const int32_t *scol  = reinterpret_cast<const int32_t *>(stage);
              const uint2 *  smeta = reinterpret_cast<const uint2 *>(stage + B::COL_BYTES);
              XT *           sxv = reinterpret_cast<XT *>(stage + B::COL_BYTES + B::META_PAD);
              const int      nblk = int(smeta[SPMV_TILE_ROWS + 1].x);
              double         r[UB][DIM];
#pragma unroll
              for (int u = 0; u < UB; ++u)
                {
                  const int j = g + u * GL;
                  if (j < nblk)
                    {
                      const int64_t col = scol[j];
#pragma unroll
                      for (int d0 = 0; d0 < DIM; ++d0)
                        r[u][d0] = __ldg(x + col * DIM + d0);
                    }
                }
#pragma unroll
              for (int u = 0; u < UB; ++u)
                {
                  const int j = g + u * GL;
                  if (j < nblk)
#pragma unroll
                    for (int d0 = 0; d0 < DIM; ++d0)
                      sxv[j * DIM + d0] = XT(r[u][d0]);
                }
              __syncwarp();
              if (lane == 0)
                mbar_arrive(smem_u32(&xfull[sx]));
            }
        }
      else
        {
          // ------------------------------ consumer warps: shared-memory FMA --------------------
          const int cw = warp - 1 - C::GATHER_WARPS;
          for (int k = 0; k < n_my; ++k)
            {
              const int sv = k % C::VSTAGES, sx = k % C::XSTAGES;
              // cfull too: the row records arrive through the async proxy, observe their barrier
              mbar_wait(smem_u32(&cfull[sx]), (k / C::XSTAGES) & 1);
              mbar_wait(smem_u32(&xfull[sx]), (k / C::XSTAGES) & 1);
              mbar_wait(smem_u32(&vfull[sv]), (k / C::VSTAGES) & 1);
              const unsigned char *xs    = xring + sx * C::XSTAGE_BYTES;
              const uint2 *        smeta = reinterpret_cast<const uint2 *>(xs + B::COL_BYTES);
              const VT *sval = reinterpret_cast<const VT *>(vring + sv * B::VAL_BYTES) +
                               (sizeof(VT) == 4 ? int(smeta[SPMV_TILE_ROWS + 1].y) : 0);
              const XT *sxv = reinterpret_cast<const XT *>(xs + B::COL_BYTES + B::META_PAD);
              const uint2 hdr    = smeta[SPMV_TILE_ROWS];
              const int   row0   = int(hdr.x);
              const int   n_rows = int(hdr.y);
              for (int row = cw; row < n_rows; row += C::CONS_WARPS)
                {
                  const uint2   m      = smeta[row];
                  const int     ne     = int(m.y >> 16) * DIM;
                  const int     stride = (ne + 1) & ~1;
                  const VT *    v      = sval + m.x;
                  const XT *    xr     = sxv + int(m.y & 0xffffu) * DIM;
                  const int64_t i      = int64_t(row0 + row) * DIM + lane;
                  if constexpr (sizeof(XT) == 4)
                    {
                      // all-FP32 row: values, x and accumulation in single precision; the .x and
                      // .y products go to separate accumulators (two independent FFMA chains)
                      float a0[DIM], a1[DIM];
#pragma unroll
                      for (int r = 0; r < DIM; ++r)
                        a0[r] = a1[r] = 0.f;
#pragma unroll 3
                      for (int e = 2 * lane; e < ne; e += 64)
                        {
                          const float x0 = xr[e];
                          const float x1 = (e + 1 < ne) ? xr[e + 1] : 0.f;
#pragma unroll
                          for (int r = 0; r < DIM; ++r)
                            {
                              const float2 vv =
                                *reinterpret_cast<const float2 *>(v + r * stride + e);
                              a0[r] = fmaf(vv.x, x0, a0[r]);
                              a1[r] = fmaf(vv.y, x1, a1[r]);
                            }
                        }
#pragma unroll
                      for (int r = 0; r < DIM; ++r)
                        {
                          float t = a0[r] + a1[r];
#pragma unroll
                          for (int o = 16; o > 0; o >>= 1)
                            t += __shfl_xor_sync(0xffffffffu, t, o);
                          a0[r] = t;
                        }
                      if (lane < DIM)
                        {
                          float yr = a0[0];
#pragma unroll
                          for (int r = 1; r < DIM; ++r)
                            if (lane == r)
                              yr = a0[r];
                          y[i] = double(yr);
                        }
                      continue;
                    }
                  // lanes that end up with a row sum: lane r (plain butterfly) or lane r*LS (TR)
                  constexpr int LS      = TR ? (DIM == 3 ? 8 : 16) : 1;
                  const int     r_mine  = lane / LS;
                  const bool    i_store = (lane % LS) == 0 && r_mine < DIM;
                  const int64_t i_mine  = int64_t(row0 + row) * DIM + r_mine;
                  double        xi      = 0.0;
                  if (DOT && i_store)
                    xi = __ldg(x + i_mine);
                  double acc[DIM];
#pragma unroll
                  for (int r = 0; r < DIM; ++r)
                    acc[r] = 0.0;
#pragma unroll 3
                  for (int e = 2 * lane; e < ne; e += 64)
                    {
                      const double x0 = xr[e];
                      const double x1 = (e + 1 < ne) ? xr[e + 1] : 0.0;
#pragma unroll
                      for (int r = 0; r < DIM; ++r)
                        {
                          const double2 vv = lds2(v + r * stride + e);
                          acc[r]           = fma(vv.x, x0, acc[r]);
                          acc[r]           = fma(vv.y, x1, acc[r]);
                        }
                    }
                  double yr;
                  if constexpr (TR)
                    yr = warp_sum_rows<DIM>(acc, lane);
                  else
                    {
#pragma unroll
                      for (int r = 0; r < DIM; ++r)
                        acc[r] = warp_sum(acc[r]);
                      yr = acc[0];
#pragma unroll
                      for (int r = 1; r < DIM; ++r)
                        if (lane == r)
                          yr = acc[r];
                    }
                  if (i_store)
                    {
                      y[i_mine] = yr;
                      if (DOT)
                        dot = fma(yr, xi, dot);
                    }
                }
              __syncwarp();
              if (lane == 0)
                {
                  mbar_arrive(smem_u32(&vempty[sv])); // release both stages
                  mbar_arrive(smem_u32(&xempty[sx]));
                }
            }
        }
      if (DOT)
        {
          const int cw = warp - 1 - C::GATHER_WARPS;
          dot          = warp_sum(dot);
          if (lane == 0 && cw >= 0)
            red[cw] = dot;
          __syncthreads();
          if (tid < 32)
            {
              double v = tid < C::CONS_WARPS ? red[tid] : 0.0;
              v        = warp_sum(v);
              if (tid == 0)
                partials[blockIdx.x] = v;
            }
        }
    }

    template <int DIM, bool DOT, typename VT, typename XT, int GW, int GG, int CW, bool TR = false>
    void launch_tma2_w(gf_context &c, const VT *val, const double *x, double *y,
                       double *dot_partials, const int *st)
    {
      using C = Tma2Cfg<DIM, VT, GW, GG, CW>;
      static bool configured = false;
      if (!configured)
        {
          GF_CUDA_CHECK(cudaFuncSetAttribute(spmv_tma2_kernel<DIM, DOT, VT, XT, GW, GG, CW, TR>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM_BYTES));
          configured = true;
        }
      const int grid = int(std::min<int64_t>(c.n_tiles, c.sm_count));
      spmv_tma2_kernel<DIM, DOT, VT, XT, GW, GG, CW, TR><<<grid, C::THREADS, C::SMEM_BYTES, c.stream>>>(
        int(c.n_tiles), c.tile_desc.p, c.tile_meta.p, c.bcol.p, val, x, y, dot_partials, st);
    }
    // two-ring kernel: 8 gather warps in 2 groups + 16 consumer warps (kinds 3 and 6). The splits
    // 8+8 and 4+16 of round 1 measured slower (0.600 / 0.572 ms against 0.544 ms,
    // profiles/r02_spmv_kernel_kinds.jsonl) and were removed.
    template <int DIM, bool DOT, typename VT>
    void launch_tma2_t(gf_context &c, int kind, const VT *val, const double *x, double *y,
                       double *dot_partials, const int *st)
    {
      if (kind == 6) // kind 3 with the transposed row reduction (the default)
        launch_tma2_w<DIM, DOT, VT, double, 8, 2, 16, true>(c, val, x, y, dot_partials, st);
      else
        launch_tma2_w<DIM, DOT, VT, double, 8, 2, 16>(c, val, x, y, dot_partials, st);
    }

    // GF_OPT_SPMV_KERNEL = 0 (default) is what has been MEASURED fastest on the B200
    // (profiles/r02_spmv_kernel_kinds.jsonl): the two-ring kernel with 16 consumer warps and the
    // transposed row reduction (kind 6) for every launch, with or without the fused dot product:
    // 0.517 ms = 6.43 TB/s on the cfg3 tangent (98 % of the measured copy peak), bitwise the y of
    // every other kind.
    int effective_kind(const gf_context &c, bool /*fused_dot*/)
    {
      return c.spmv_kernel_kind == 0 ? 6 : c.spmv_kernel_kind;
    }

    template <int DIM, bool DOT, typename VT>
    void launch_tma_t(gf_context &c, const VT *val, const double *x, double *y,
                      double *dot_partials, const int *st)
    {
      using C = TmaCfg<DIM, VT>;
      static bool configured = false;
      if (!configured)
        {
          GF_CUDA_CHECK(cudaFuncSetAttribute(spmv_tma_kernel<DIM, DOT, VT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM_BYTES));
          configured = true;
        }
      const int grid = int(std::min<int64_t>(c.n_tiles, c.sm_count));
      spmv_tma_kernel<DIM, DOT, VT><<<grid, C::THREADS, C::SMEM_BYTES, c.stream>>>(
        int(c.n_tiles), c.tile_desc.p, c.tile_meta.p, c.bcol.p, val, x, y, dot_partials, st);
    }

#else  // GF_CUDA_EMULATION: the CPU stand-in of tests/cuda_emu has no TMA; every launch takes
       // the LDG kernel (kind 1), the tile lists are still built (pattern.cu)
    int effective_kind(const gf_context &, bool) { return 1; }
    template <int DIM, bool DOT, typename VT>
    void launch_tma2_t(gf_context &, int, const VT *, const double *, double *, double *, const int *)
    {
      throw Error{GF_ERR_UNSUPPORTED, "TMA SpMV kernels do not exist in the CPU emulation"};
    }
    template <int DIM, bool DOT, typename VT>
    void launch_tma_t(gf_context &, const VT *, const double *, double *, double *, const int *)
    {
      throw Error{GF_ERR_UNSUPPORTED, "TMA SpMV kernels do not exist in the CPU emulation"};
    }
    template <int DIM, bool DOT, typename VT, typename XT, int GW, int GG, int CW, bool TR = false>
    void launch_tma2_w(gf_context &, const VT *, const double *, double *, double *, const int *)
    {
      throw Error{GF_ERR_UNSUPPORTED, "TMA SpMV kernels do not exist in the CPU emulation"};
    }
#endif // GF_CUDA_EMULATION

    // FP64 array -> FP32 copy with the same indexing (n even: every scalar row is padded)
    __global__ void convert_f32_kernel(const int64_t n_pairs, const double2 *__restrict__ in,
                                       float2 *__restrict__ out)
    {
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n_pairs;
           i += int64_t(gridDim.x) * blockDim.x)
        {
          const double2 v = __ldcs(in + i);
          out[i]          = make_float2(float(v.x), float(v.y));
        }
    }

    // y = M x with M = m_ab delta_cd (consistent mass, one scalar per block)
    template <int DIM>
    __global__ void spmv_mass_kernel(const int64_t n_rows, const int32_t *__restrict__ brow_ptr,
                                     const int32_t *__restrict__ bcol,
                                     const double *__restrict__ mass_blk,
                                     const double *__restrict__ x, double *__restrict__ y)
    {
      const int     lane = threadIdx.x & 31;
      const int64_t A    = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
      if (A >= n_rows)
        return;
      const int32_t b0 = brow_ptr[A], b1 = brow_ptr[A + 1];
      double        acc[DIM];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        acc[r] = 0.0;
      for (int32_t k = b0 + lane; k < b1; k += 32)
        {
          const double  m = mass_blk[k];
          const int64_t B = bcol[k];
#pragma unroll
          for (int r = 0; r < DIM; ++r)
            acc[r] = fma(m, x[B * DIM + r], acc[r]);
        }
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        acc[r] = warp_sum(acc[r]);
      if (lane == 0)
#pragma unroll
        for (int r = 0; r < DIM; ++r)
          y[A * DIM + r] = acc[r];
    }
  } // namespace

  // number of per-CTA partial sums the fused dot product of launch_spmv writes
  int spmv_dot_partials(const gf_context &c)
  {
    if (c.n_tiles > 0 && c.spmv_kernel_kind != 1)
      return int(std::min<int64_t>(c.n_tiles, c.sm_count));
    const int64_t want = (c.n_owned_nodes * 32 + SPMV_THREADS - 1) / SPMV_THREADS;
    return int(std::min<int64_t>(want, c.max_red_blocks));
  }

  void launch_spmv(gf_context &c, const double *val, const double *x, double *y,
                   double *dot_partials)
  {
    ProfScope     ps(c, c.mg_level > 0 ? Profile::MG_SPMV : Profile::SPMV);
    const int64_t n_rows = c.n_owned_nodes;
    if (n_rows == 0)
      return;
    // the CG's dot products are chunked reductions (reduce.cuh) since round 2: the fused variants
    // summed per CTA, i.e. partition dependent, and are no longer instantiated
    GF_REQUIRE(dot_partials == nullptr, GF_ERR_INVALID_ARG, "fused dot product is not available");
    const int kind = effective_kind(c, false);
    if (c.n_tiles > 0 && (kind == 3 || kind == 6))
      {
        if (c.dim == 3)
          launch_tma2_t<3, false, double>(c, kind, val, x, y, nullptr, nullptr);
        else
          launch_tma2_t<2, false, double>(c, kind, val, x, y, nullptr, nullptr);
        GF_CUDA_CHECK(cudaGetLastError());
        return;
      }
    if (c.n_tiles > 0 && kind == 5)
      {
        if (c.dim == 3)
          launch_tma_t<3, false, double>(c, val, x, y, nullptr, nullptr);
        else
          launch_tma_t<2, false, double>(c, val, x, y, nullptr, nullptr);
        GF_CUDA_CHECK(cudaGetLastError());
        return;
      }
    const int grid = spmv_dot_partials(c);
    if (c.dim == 3)
      spmv_kernel<3, false, double><<<grid, SPMV_THREADS, 0, c.stream>>>(
        n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val, x, y, nullptr, nullptr);
    else
      spmv_kernel<2, false, double><<<grid, SPMV_THREADS, 0, c.stream>>>(
        n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val, x, y, nullptr, nullptr);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  // y = A32 x: the FP32 copy of a level operator inside the multigrid V-cycle (no fused dot)
  void launch_spmv_f32(gf_context &c, const float *val32, const double *x, double *y)
  {
    ProfScope     ps(c, c.mg_level > 0 ? Profile::MG_SPMV : Profile::SPMV);
    const int64_t n_rows = c.n_owned_nodes;
    if (n_rows == 0)
      return;
    const int kind = effective_kind(c, false);
    if (c.n_tiles > 0 && (kind == 3 || kind == 6))
      {
        if (c.dim == 3)
          launch_tma2_t<3, false, float>(c, kind, val32, x, y, nullptr, nullptr);
        else
          launch_tma2_t<2, false, float>(c, kind, val32, x, y, nullptr, nullptr);
      }
    else if (c.n_tiles > 0 && kind == 5)
      {
        if (c.dim == 3)
          launch_tma_t<3, false, float>(c, val32, x, y, nullptr, nullptr);
        else
          launch_tma_t<2, false, float>(c, val32, x, y, nullptr, nullptr);
      }
    else
      {
        const int grid = spmv_dot_partials(c);
        if (c.dim == 3)
          spmv_kernel<3, false, float><<<grid, SPMV_THREADS, 0, c.stream>>>(
            n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val32, x, y, nullptr, nullptr);
        else
          spmv_kernel<2, false, float><<<grid, SPMV_THREADS, 0, c.stream>>>(
            n_rows, c.brow_ptr.p, c.val_ptr.p, c.bcol.p, val32, x, y, nullptr, nullptr);
      }
    GF_CUDA_CHECK(cudaGetLastError());
  }

  // y = A32 x with x staged and accumulated in FP32 as well (GF_OPT_MG_MATRIX_PRECISION = 2);
  // matrices without tiles fall back to the FP32-value / FP64-accumulate kernels
  void launch_spmv_f32x(gf_context &c, const float *val32, const double *x, double *y)
  {
    if (c.n_tiles == 0 || effective_kind(c, false) == 1 || c.n_owned_nodes == 0)
      {
        launch_spmv_f32(c, val32, x, y);
        return;
      }
    ProfScope ps(c, c.mg_level > 0 ? Profile::MG_SPMV : Profile::SPMV);
    if (c.dim == 3)
      launch_tma2_w<3, false, float, float, 8, 2, 16>(c, val32, x, y, nullptr, nullptr);
    else
      launch_tma2_w<2, false, float, float, 8, 2, 16>(c, val32, x, y, nullptr, nullptr);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_convert_f32(gf_context &c, const double *val, float *val32)
  {
    ProfScope     ps(c, Profile::MG_VEC);
    const int64_t n_pairs = c.n_val / 2; // n_val is even (padded scalar rows)
    if (n_pairs == 0)
      return;
    const int grid = int(std::min<int64_t>((n_pairs + 255) / 256, int64_t(c.sm_count) * 16));
    convert_f32_kernel<<<grid, 256, 0, c.stream>>>(
      n_pairs, reinterpret_cast<const double2 *>(val), reinterpret_cast<float2 *>(val32));
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void launch_spmv_mass(gf_context &c, const double *x, double *y)
  {
    ProfScope     ps(c, Profile::SPMV);
    const int64_t n_rows = c.n_owned_nodes;
    if (n_rows == 0)
      return;
    const unsigned grid = unsigned((n_rows * 32 + 255) / 256);
    if (c.dim == 3)
      spmv_mass_kernel<3><<<grid, 256, 0, c.stream>>>(n_rows, c.brow_ptr.p, c.bcol.p,
                                                      c.mass_blk.p, x, y);
    else
      spmv_mass_kernel<2><<<grid, 256, 0, c.stream>>>(n_rows, c.brow_ptr.p, c.bcol.p,
                                                      c.mass_blk.p, x, y);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  // bytes one SpMV launch must move in the stored format (the roofline numerator)
  double spmv_bytes(const gf_context &c)
  {
    // values + column indices + per-row records (12 B/row in both kernels: row pointers for the LDG
    // kernel, tile records for the TMA kernel are of the same order) + x read once + y written once
    return 8.0 * double(c.n_val) + 4.0 * double(c.n_blocks) + 12.0 * double(c.n_owned_nodes + 1) +
           8.0 * double(c.n_local) + 8.0 * double(c.n_owned);
  }
  // the same with the FP32 value copy (vectors stay FP64)
  double spmv_bytes_f32(const gf_context &c)
  {
    return spmv_bytes(c) - 4.0 * double(c.n_val);
  }
} // namespace gf
