// Split-precision (bf16 x 3) tensor-core GEMM / implicit 5x5 convolution for sm_100a.
//
//   C[m, n] = act(alpha * sum_{tap, k} A[m + off(tap), k] * B[tap][n][k] + bias[n]) + beta * R[m, n]
//
// fp32 operands are pre-split by `dfold_split2d` / `dfold_conv_weight_prep` into bf16 planes
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi)); the kernel accumulates hi*hi + hi*lo + lo*hi in fp32 in
// TMEM, which keeps ~16 mantissa bits per operand (relative error ~2^-17 per product, vs 2^-11 for plain
// TF32/BF16 — SURVEY.md §0 "precision trap").
//
// Structure (one CTA = one 128 x BN output tile, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.3d (128B swizzle) of the A / B hi+lo tiles into a
//               multi-stage smem ring; the 3-D tensor map of the activations makes the 5x5 halo an
//               out-of-bounds zero fill, so the convolution needs no im2col and no padding buffer.
//   warp 1      tcgen05.mma issuer (one elected lane): 3 products x 4 K-steps of UMMA 128xBNx16 per stage,
//               accumulator in TMEM; tcgen05.commit releases the smem stage / signals the epilogue.
//   warps 2..9  accumulate + epilogue.  The tensor core's fp32 accumulator TRUNCATES on every MMA (measured on
//               B200: relative bias -1e-4 at K = 32000, growing linearly with the number of MMAs), so a TMEM
//               accumulator only ever holds kChunk = 4 k-blocks (48 MMAs).  Two TMEM accumulators ping-pong: while
//               the MMA warp fills one, these warps tcgen05.ld the other and add it into fp32 registers with
//               round-to-nearest CUDA-core adds, then apply bias / activation / residual and store.
//
// Replaces (reference file:line): nn.Linear in src/model/ipa_pytorch_dynamic.py:284-305,757-796,590,
// openfold/model/structure_module.py:102-110,58-59; nn.Conv2d stack src/model/ipa_pytorch_dynamic.py:664-706.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace dfold {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 320;     // 2 control warps + 8 accumulate/epilogue warps
constexpr int kChunk = 4;         // k-blocks accumulated in TMEM before promotion to registers

template <int BN> struct TileCfg {
    static constexpr int kStages = (BN >= 256) ? 2 : ((BN >= 128) ? 3 : 4);
    static constexpr int kABytes = BM * BK * 2;               // one plane
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int kTmemCols = 2 * BN;                  // two ping-pong accumulators (power of two >= 32)
};

struct GemmParams {
    int mode;            // 0: A rows = (frame, residue) tiles, taps outer / K inner   1: wgrad (K = pixels)
    int num_kb;          // k-blocks per tile
    int kc;              // mode 0: k-chunks per tap; mode 1: residue blocks per frame
    int taps_n, taps_f;  // tap grid (1x1 for a plain linear)
    int tiles_per_frame; // mode 0: ceil(Nr / 128)
    int Nr;              // mode 0: residues per frame (rows per frame)
    long out_rows;       // valid rows of the output (mode 0: F*Nr, mode 1: M)
    int n_out;           // valid columns
    float* out; long ldo; long out_tap_stride;
    const float* bias;
    const float* res; long ldr;
    float alpha, beta;
    int act;             // 0 none, 1 relu, 2 silu
    // ---- batching (defaults describe the plain conv / linear case) ----
    // mode 0: the "frame" index f0 of a row tile doubles as a batch id: hb = f0 % bmod
    int bmod;            // >= 1
    int a_f_div;         // A frame coordinate            = f0 / a_f_div
    int a_k_bstride;     // A K-coordinate offset         = hb * a_k_bstride
    int b_k_ofs;         // B K-coordinate offset         = b_k_ofs + hb * b_k_bstride
    int b_k_bstride;
    int b_z_bstride;     // B third coordinate            = tap + hb * b_z_bstride
    int f_start;         // A frame coordinate offset (output frame f reads input frames f + f_start + df): cropped convs
    int b_f_add;         // mode 1: B outer coordinate offset
    int o_f_div;         // output row                    = (f0 / o_f_div) * Nr + n
    long o_col_bstride;  // output column offset          = hb * o_col_bstride
    // mode 1: blockIdx.z = zq * zdiv + zs  (zq: tap or batch id, zs: K split)
    int zdiv;
    int a_f_mul, a_z_mul;   // A outer coordinate = f * a_f_mul + zq * a_z_mul
    int b_f_mul, b_z_mul;   // B outer coordinate = f * b_f_mul + zq * b_z_mul (+ df)
    int b_n_zmul;           // B column offset    = zq * b_n_zmul
    int kb_per_split;       // k-blocks per K split
    int atomic;             // epilogue accumulates with atomicAdd (split-K)
    int shift_on_a;         // pair kernel, mode 1: the tap shift / frame offset applies to operand A (roles swapped)
    long o_rs, o_cs;        // pair kernel, mode 1: output element (row m, column n) at out[m * o_rs + n * o_cs]
    int raster_n;           // pair kernel, mode 0: column tiles fastest in the dispatch order
    long long* stats;       // debug (dfold_debug_gemm_stats): per CTA {total, MMA wait full, MMA wait acc_empty, acc warp wait} cycles
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset   [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

// MN-major, 128B-swizzled operand tile (weight-gradient mode): the tile is stored [k][mn] with 64 mn-elements
// (128 B) contiguous per k row; 8 k-rows form a 1024 B swizzle atom (SBO), 64-wide mn atoms are 64 rows x 128 B
// = 8192 B apart (LBO).
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(8192 >> 4) << 16;                  // leading byte offset: next 64-wide mn atom
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next group of 8 k rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const GemmParams p) {
    using Cfg = TileCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;   // 8-byte aligned
    // barrier layout: full[kStages], empty[kStages], acc_full[2], acc_empty[2], then the tmem pointer word
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto acc_full_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + b); };
    auto acc_empty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- tile coordinates ----
    const int n_tile = blockIdx.x;                 // output column tile
    const int m_tile = blockIdx.y;
    const int zq = (p.mode == 1) ? (int)blockIdx.z / p.zdiv : 0;     // mode 1: tap or batch id
    const int zs = (p.mode == 1) ? (int)blockIdx.z % p.zdiv : 0;     // mode 1: K split
    int f0 = 0, n0 = 0, m0 = 0, hb = 0;
    int kb_begin = 0, num_kb = p.num_kb;
    if (p.mode == 0) {
        f0 = m_tile / p.tiles_per_frame;
        n0 = (m_tile % p.tiles_per_frame) * BM;
        hb = f0 % p.bmod;
        kb_begin = (int)blockIdx.z * p.kb_per_split;          // split-K over (tap, k-chunk) for launches that cannot fill the chip
        num_kb = min(p.kb_per_split, p.num_kb - kb_begin);
    } else {
        m0 = m_tile * BM;
        kb_begin = zs * p.kb_per_split;
        num_kb = min(p.kb_per_split, p.num_kb - kb_begin);
    }
    const int col0 = n_tile * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full_bar(b), 1); mbar_init(acc_empty_bar(b), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
                const uint32_t sa_hi = smem_base + s * Cfg::kStageBytes;
                const uint32_t sa_lo = sa_hi + Cfg::kABytes;
                const uint32_t sb_hi = sa_lo + Cfg::kABytes;
                const uint32_t sb_lo = sb_hi + Cfg::kBBytes;
                if (p.mode == 0) {
                    const int tap = (kb_begin + kb) / p.kc, c = (kb_begin + kb) % p.kc;
                    const int dn = tap % p.taps_n - p.taps_n / 2;
                    const int df = tap / p.taps_n - p.taps_f / 2;
                    const int ak = c * BK + hb * p.a_k_bstride;
                    const int af = f0 / p.a_f_div + p.f_start + df;
                    const int bk = c * BK + p.b_k_ofs + hb * p.b_k_bstride;
                    const int bz = tap + hb * p.b_z_bstride;
                    tma_load_3d(sa_hi, &map_a_hi, full_bar(s), ak, n0 + dn, af);
                    tma_load_3d(sa_lo, &map_a_lo, full_bar(s), ak, n0 + dn, af);
                    tma_load_3d(sb_hi, &map_b_hi, full_bar(s), bk, col0, bz);
                    tma_load_3d(sb_lo, &map_b_lo, full_bar(s), bk, col0, bz);
                } else {
                    // K runs over pixels (f, 64-residue block j); operands are MN-major: boxes of
                    // 64 channels x 64 residues, one per 64-wide channel atom.  The tap shift is on the
                    // residue / frame coordinates (row dimensions), out-of-image rows are zero-filled.
                    const int kg = kb_begin + kb;
                    const int f = kg / p.kc, j = kg % p.kc;
                    const bool has_taps = p.taps_n * p.taps_f > 1;
                    const int dn = has_taps ? zq % p.taps_n - p.taps_n / 2 : 0;
                    const int df = has_taps ? zq / p.taps_n - p.taps_f / 2 : 0;
                    const int af = f * p.a_f_mul + zq * p.a_z_mul;
                    const int bf = f * p.b_f_mul + zq * p.b_z_mul + p.b_f_add + df;
                    const int bn0 = col0 + zq * p.b_n_zmul;
#pragma unroll
                    for (int a = 0; a < BM / 64; ++a) {
                        tma_load_3d(sa_hi + a * 8192, &map_a_hi, full_bar(s), m0 + a * 64, j * BK, af);
                        tma_load_3d(sa_lo + a * 8192, &map_a_lo, full_bar(s), m0 + a * 64, j * BK, af);
                    }
#pragma unroll
                    for (int b = 0; b < BN / 64; ++b) {
                        tma_load_3d(sb_hi + b * 8192, &map_b_hi, full_bar(s), bn0 + b * 64, j * BK + dn, bf);
                        tma_load_3d(sb_lo + b * 8192, &map_b_lo, full_bar(s), bn0 + b * 64, j * BK + dn, bf);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
            uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const bool mn_major = (p.mode == 1);
            if (mn_major) idesc |= (1u << 15) | (1u << 16);     // A and B are MN-major in weight-gradient mode
            long long w_full = 0, w_acc = 0, t_begin = 0;
            if (p.stats) t_begin = clock64();
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                const int chunk = kb / kChunk;
                const int buf = chunk & 1;
                const bool chunk_start = (kb % kChunk) == 0;
                if (chunk_start) {
                    // the accumulate warps must have drained this TMEM buffer (two chunks ago)
                    const long long t0 = p.stats ? clock64() : 0;
                    mbar_wait(acc_empty_bar(buf), ((chunk >> 1) & 1) ^ 1u);
                    if (p.stats) w_acc += clock64() - t0;
                    tcgen05_fence_after();
                }
                {
                    const long long t0 = p.stats ? clock64() : 0;
                    mbar_wait(full_bar(s), ph);
                    if (p.stats) w_full += clock64() - t0;
                }
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                const uint32_t sa_hi = smem_base + s * Cfg::kStageBytes;
                const uint32_t sa_lo = sa_hi + Cfg::kABytes;
                const uint32_t sb_hi = sa_lo + Cfg::kABytes;
                const uint32_t sb_lo = sb_hi + Cfg::kBBytes;
                const uint64_t da_hi = mn_major ? make_sw128_mn_desc(sa_hi) : make_sw128_desc(sa_hi);
                const uint64_t da_lo = mn_major ? make_sw128_mn_desc(sa_lo) : make_sw128_desc(sa_lo);
                const uint64_t db_hi = mn_major ? make_sw128_mn_desc(sb_hi) : make_sw128_desc(sb_hi);
                const uint64_t db_lo = mn_major ? make_sw128_mn_desc(sb_lo) : make_sw128_desc(sb_lo);
                // per K step (16 elements): K-major +32 B inside the swizzle row; MN-major +2 groups of 8 k rows
                const uint64_t kstep = mn_major ? (uint64_t)((2 * 1024) >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t koff = kstep * k;
                    umma_bf16(tmem_d, da_lo + koff, db_hi + koff, idesc, (!chunk_start || k > 0) ? 1u : 0u);
                    umma_bf16(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                    umma_bf16(tmem_d, da_hi + koff, db_hi + koff, idesc, 1u);
                }
                umma_commit(empty_bar(s));      // smem stage reusable once these MMAs retire
                if ((kb % kChunk) == kChunk - 1 || kb == num_kb - 1)
                    umma_commit(acc_full_bar(buf));   // this chunk's partial sum is complete
            }
            if (p.stats) {
                const long cta = ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
                p.stats[4 * cta + 0] = clock64() - t_begin;
                p.stats[4 * cta + 1] = w_full;
                p.stats[4 * cta + 2] = w_acc;
            }
        }
    } else {
        // =============================== accumulate + epilogue (warps 2..9) ===============================
        constexpr int HALF = BN / 2;                    // columns owned by this warp
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which half of the tile's columns
        float acc[HALF];
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
        const int nchunks = (num_kb + kChunk - 1) / kChunk;
        long long w_accfull = 0;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int buf = chunk & 1;
            const long long t0 = p.stats ? clock64() : 0;
            mbar_wait(acc_full_bar(buf), (chunk >> 1) & 1);
            if (p.stats) w_accfull += clock64() - t0;
            tcgen05_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * HALF + c0), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[c0 + i] += __uint_as_float(v[i]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar(buf));
        }
        if (p.stats && warp == 2 && lane == 0) {
            const long cta = ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            p.stats[4 * cta + 3] = w_accfull;
        }
        const int r = q * 32 + lane;                    // row inside the tile
        long grow;                                      // global output row
        bool row_ok;
        if (p.mode == 0) {
            const int n = n0 + r;
            grow = (long)(f0 / p.o_f_div) * p.Nr + n;
            row_ok = (n < p.Nr) && (grow < p.out_rows);
        } else {
            grow = m0 + r;
            row_ok = grow < p.out_rows;
        }
        const long obase = (p.mode == 1 ? (long)zq * p.out_tap_stride : (long)hb * p.o_col_bstride);
        float* orow = p.out + obase + grow * p.ldo;
        const float* rrow = p.res ? p.res + (p.mode == 0 ? (long)hb * p.o_col_bstride : 0) + grow * p.ldr : nullptr;
        const bool vec_ok = !p.atomic && ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                            (p.out_tap_stride % 4 == 0) && (p.o_col_bstride % 4 == 0);
        if (row_ok) {
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 4) {
                const int gc0 = col0 + half * HALF + c0;
                if (gc0 >= p.n_out) break;
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float x = acc[c0 + i] * p.alpha;
                    const int gc = gc0 + i;
                    if (gc < p.n_out) {
                        if (p.bias) x += __ldg(p.bias + gc);
                        if (p.act == 1) x = fmaxf(x, 0.f);
                        else if (p.act == 2) x = x / (1.f + __expf(-x));
                        if (rrow) x += p.beta * __ldg(rrow + gc);
                    }
                    o[i] = x;
                }
                if (vec_ok && gc0 + 4 <= p.n_out) {
                    *reinterpret_cast<float4*>(orow + gc0) = make_float4(o[0], o[1], o[2], o[3]);
                } else if (p.atomic) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gc0 + i < p.n_out) atomicAdd(orow + gc0 + i, o[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gc0 + i < p.n_out) orow[gc0 + i] = o[i];
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols));
    }
}


// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster compute a 256 x BN tile.  Each CTA stages its own 128 rows of
// A and HALF of the B tile (BN/2 rows) and the pair's tensor cores read both halves, so per SM the operand traffic per
// MMA cycle halves compared with the single-CTA 128 x 128 tile: measured with dfold_debug_gemm_stats the single-CTA
// kernel's MMA thread never waits on data yet issues at 65 % of the MMA floor — shared-memory bandwidth (operand reads
// 128 B/clk + TMA writes 85 B/clk per SM) is the limiter.  Here: reads 64 B/clk + writes 42 B/clk (BN = 256).
//   full[s]      lives in the LEADER CTA (rank 0): count 2 (one arrival per CTA's producer) + the bytes of both CTAs
//   empty[s]     per CTA, released by the leader's MMA thread with a multicast tcgen05.commit
//   acc_full[b]  per CTA (multicast commit); acc_empty[b] in the leader, count 16 (8 accumulate warps x 2 CTAs)
// K-major operands only (forward / data-gradient / linear), no batching.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address -> rank 0

template <int BN> struct PairCfg {
    static constexpr int kBRows = BN / 2;                      // B rows staged per CTA
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = kBRows * BK * 2;
    static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr int kStages = (212 * 1024) / kStageBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
    static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
    static_assert(kBBytes % 1024 == 0, "B half tile must be a whole number of swizzle atoms");
    static_assert(BN % 32 == 0 && BN <= 256, "pair tile width");
};

__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_leader, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar_leader) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const GemmParams p) {
    using Cfg = PairCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto acc_full_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + b); };
    auto acc_empty_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;

    // ---- tile coordinates: the pair owns row tiles 2j, 2j+1 and one BN-wide column tile ----
    // Raster order: the hardware dispatches CTAs x-fastest.  With `raster_n` the linear pair index runs over the column
    // tiles first, so the pairs in flight share a narrow band of activation rows and the whole weight tensor (both L2
    // resident) instead of streaming the entire activation once per column tile.
    int m_tile = blockIdx.x;
    int n_tile = blockIdx.y;
    if (p.raster_n) {
        const int lin = (int)blockIdx.y * ((int)gridDim.x >> 1) + ((int)blockIdx.x >> 1);
        n_tile = lin % (int)gridDim.y;
        m_tile = 2 * (lin / (int)gridDim.y) + ((int)blockIdx.x & 1);
    }
    const int zq = blockIdx.z;                                 // mode 1: tap
    const int f0 = m_tile / p.tiles_per_frame;                 // mode 0; beyond the last frame for a padding CTA: loads zero-fill
    const int n0 = (m_tile % p.tiles_per_frame) * BM;
    const int m0 = m_tile * BM;                                // mode 1
    const int col0 = n_tile * BN;
    const int num_kb = p.num_kb;
    const bool mn_major = p.mode == 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full_bar(b), 1); mbar_init(acc_empty_bar(b), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // =============================== TMA producer (both CTAs) ===============================
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(empty_bar(s), ph ^ 1u);
                const uint32_t fb = full_bar(s) & kPeerMask;           // the leader's barrier
                if (leader) mbar_expect_tx(full_bar(s), 2 * Cfg::kStageBytes);
                const uint32_t sa_hi = smem_base + s * Cfg::kStageBytes;
                const uint32_t sa_lo = sa_hi + Cfg::kABytes;
                const uint32_t sb_hi = sa_lo + Cfg::kABytes;
                const uint32_t sb_lo = sb_hi + Cfg::kBBytes;
                if (!mn_major) {
                    const int tap = kb / p.kc, c = kb % p.kc;
                    const int dn = tap % p.taps_n - p.taps_n / 2;
                    const int df = tap / p.taps_n - p.taps_f / 2;
                    const int ak = c * BK;
                    const int af = f0 + p.f_start + df;
                    const int brow = col0 + (int)rank * Cfg::kBRows;
                    tma_load_3d_pair(sa_hi, &map_a_hi, fb, ak, n0 + dn, af);
                    tma_load_3d_pair(sa_lo, &map_a_lo, fb, ak, n0 + dn, af);
                    tma_load_3d_pair(sb_hi, &map_b_hi, fb, ak, brow, tap);
                    tma_load_3d_pair(sb_lo, &map_b_lo, fb, ak, brow, tap);
                } else {
                    // K = pixels (frame f, 64-residue block j); MN-major boxes of 64 channels x 64 residues.  The tap's
                    // residue / frame shift sits on one operand's row coordinates (out-of-image rows zero-fill).
                    const int f = kb / p.kc, j = kb % p.kc;
                    const int dn = zq % p.taps_n - p.taps_n / 2;
                    const int df = zq / p.taps_n - p.taps_f / 2;
                    const int a_n = j * BK + (p.shift_on_a ? dn : 0), a_f = f + (p.shift_on_a ? p.b_f_add + df : 0);
                    const int b_n = j * BK + (p.shift_on_a ? 0 : dn), b_f = f + (p.shift_on_a ? 0 : p.b_f_add + df);
                    const int bn0 = col0 + (int)rank * Cfg::kBRows;
#pragma unroll
                    for (int a = 0; a < BM / 64; ++a) {
                        tma_load_3d_pair(sa_hi + a * 8192, &map_a_hi, fb, m0 + a * 64, a_n, a_f);
                        tma_load_3d_pair(sa_lo + a * 8192, &map_a_lo, fb, m0 + a * 64, a_n, a_f);
                    }
#pragma unroll
                    for (int b = 0; b < Cfg::kBRows / 64; ++b) {
                        tma_load_3d_pair(sb_hi + b * 8192, &map_b_hi, fb, bn0 + b * 64, b_n, b_f);
                        tma_load_3d_pair(sb_lo + b * 8192, &map_b_lo, fb, bn0 + b * 64, b_n, b_f);
                    }
                }
                if (!leader) mbar_arrive_leader(full_bar(s));
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer (leader CTA only) ===============================
        if (leader && lane == 0) {
            // D=f32, A=B=bf16, N=BN, M=256 across the pair; K-major operands, or MN-major in weight-gradient mode
            uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            if (mn_major) idesc |= (1u << 15) | (1u << 16);
            long long w_full = 0, w_acc = 0, t_begin = 0;
            if (p.stats) t_begin = clock64();
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                const int chunk = kb / kChunk;
                const int buf = chunk & 1;
                const bool chunk_start = (kb % kChunk) == 0;
                if (chunk_start) {
                    const long long t0 = p.stats ? clock64() : 0;
                    mbar_wait(acc_empty_bar(buf), ((chunk >> 1) & 1) ^ 1u);
                    if (p.stats) w_acc += clock64() - t0;
                    tcgen05_fence_after();
                }
                {
                    const long long t0 = p.stats ? clock64() : 0;
                    mbar_wait(full_bar(s), ph);
                    if (p.stats) w_full += clock64() - t0;
                }
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                const uint32_t sa_hi = smem_base + s * Cfg::kStageBytes;
                const uint32_t sa_lo = sa_hi + Cfg::kABytes;
                const uint32_t sb_hi = sa_lo + Cfg::kABytes;
                const uint32_t sb_lo = sb_hi + Cfg::kBBytes;
                const uint64_t da_hi = mn_major ? make_sw128_mn_desc(sa_hi) : make_sw128_desc(sa_hi);
                const uint64_t da_lo = mn_major ? make_sw128_mn_desc(sa_lo) : make_sw128_desc(sa_lo);
                const uint64_t db_hi = mn_major ? make_sw128_mn_desc(sb_hi) : make_sw128_desc(sb_hi);
                const uint64_t db_lo = mn_major ? make_sw128_mn_desc(sb_lo) : make_sw128_desc(sb_lo);
                const uint64_t kstep = mn_major ? (uint64_t)((2 * 1024) >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t koff = kstep * k;
                    umma_bf16_pair(tmem_d, da_lo + koff, db_hi + koff, idesc, (!chunk_start || k > 0) ? 1u : 0u);
                    umma_bf16_pair(tmem_d, da_hi + koff, db_lo + koff, idesc, 1u);
                    umma_bf16_pair(tmem_d, da_hi + koff, db_hi + koff, idesc, 1u);
                }
                umma_commit_pair(empty_bar(s));
                if ((kb % kChunk) == kChunk - 1 || kb == num_kb - 1) umma_commit_pair(acc_full_bar(buf));
            }
            if (p.stats) {
                const long cta = ((long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
                p.stats[4 * cta + 0] = clock64() - t_begin;
                p.stats[4 * cta + 1] = w_full;
                p.stats[4 * cta + 2] = w_acc;
            }
        }
    } else {
        // =============================== accumulate + epilogue (warps 2..9, both CTAs) ===============================
        constexpr int HALF = BN / 2;
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        float acc[HALF];
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
        const int nchunks = (num_kb + kChunk - 1) / kChunk;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(acc_full_bar(buf), (chunk >> 1) & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * HALF + c0), v);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[c0 + i] += __uint_as_float(v[i]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(acc_empty_bar(buf));
        }
        const int r = q * 32 + lane;
        if (mn_major) {
            // weight gradient: row = channel m0 + r; element (m, n) at out[zq * tap_stride + m * o_rs + n * o_cs]
            // (o_rs = 1 for swapped operand roles: the 32 lanes of a warp then write 32 consecutive floats)
            const long gm = (long)m0 + r;
            if (gm < p.out_rows) {
                float* ob = p.out + (long)zq * p.out_tap_stride + gm * p.o_rs;
                const bool v4 = p.o_cs == 1 && ((p.o_rs & 3) == 0) && ((p.out_tap_stride & 3) == 0) &&
                                ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
                if (v4) {
#pragma unroll
                    for (int c = 0; c < HALF; c += 4) {
                        const int gc = col0 + half * HALF + c;
                        if (gc + 4 <= p.n_out)
                            *reinterpret_cast<float4*>(ob + gc) = make_float4(acc[c] * p.alpha, acc[c + 1] * p.alpha, acc[c + 2] * p.alpha, acc[c + 3] * p.alpha);
                        else
                            for (int i = 0; i < 4; ++i)
                                if (gc + i < p.n_out) ob[gc + i] = acc[c + i] * p.alpha;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < HALF; ++c) {
                        const int gc = col0 + half * HALF + c;
                        if (gc < p.n_out) ob[(long)gc * p.o_cs] = acc[c] * p.alpha;
                    }
                }
            }
        }
        const int n = n0 + r;
        const long grow = (long)f0 * p.Nr + n;
        const bool row_ok = !mn_major && (n < p.Nr) && (grow < p.out_rows);
        float* orow = p.out + grow * p.ldo;
        const float* rrow = p.res ? p.res + grow * p.ldr : nullptr;
        const bool vec_ok = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
        if (row_ok) {
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 4) {
                const int gc0 = col0 + half * HALF + c0;
                if (gc0 >= p.n_out) break;
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float x = acc[c0 + i] * p.alpha;
                    const int gc = gc0 + i;
                    if (gc < p.n_out) {
                        if (p.bias) x += __ldg(p.bias + gc);
                        if (p.act == 1) x = fmaxf(x, 0.f);
                        else if (p.act == 2) x = x / (1.f + __expf(-x));
                        if (rrow) x += p.beta * __ldg(rrow + gc);
                    }
                    o[i] = x;
                }
                if (vec_ok && gc0 + 4 <= p.n_out) {
                    *reinterpret_cast<float4*>(orow + gc0) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gc0 + i < p.n_out) orow[gc0 + i] = o[i];
                }
            }
        }
        tcgen05_fence_before();
    }
    // the peer's tensor core reads this CTA's B half and (leader) both CTAs' barriers: leave together
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols));
    }
}

// out = act(out + bias) + beta * res   (second pass of a split-K launch; rows x n_out, row strides ldo / ldr)
__global__ void gemm_finish_kernel(float* __restrict__ out, long ldo, const float* __restrict__ bias, const float* __restrict__ res,
                                   long ldr, float beta, int act, long rows, int n_out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * n_out) return;
    const long r = i / n_out;
    const int c = (int)(i % n_out);
    float x = out[r * ldo + c];
    if (bias) x += __ldg(bias + c);
    if (act == 1) x = fmaxf(x, 0.f);
    else if (act == 2) x = x / (1.f + __expf(-x));
    if (res) x += beta * __ldg(res + r * ldr + c);
    out[r * ldo + c] = x;
}

// ---------------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 tensor with dims d0 (innermost, contiguous), d1, d2; strides in ELEMENTS for d1, d2.
int make_map(CUtensorMap* m, const void* base, long d0, long d1, long d2, long s1, long s2, int b0, int b1, int b2) {
    EncodeTiledFn enc = get_encode();
    DFOLD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    DFOLD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    DFOLD_REQUIRE((s1 * 2) % 16 == 0 && (s2 * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes (s1=%ld s2=%ld)", s1, s2);
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DFOLD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) dims=(%ld,%ld,%ld) strides=(%ld,%ld) box=(%d,%d,%d)",
                  (int)r, d0, d1, d2, s1, s2, b0, b1, b2);
    return 0;
}

long long* g_stats = nullptr;      // set by dfold_debug_gemm_stats

template <int BN>
int launch(const CUtensorMap* maps, const GemmParams& p_in, dim3 grid, cudaStream_t st) {
    using Cfg = TileCfg<BN>;
    GemmParams p = p_in;
    p.stats = g_stats;
    static SmemCfg cfg;
    if (ensure_dyn_smem(gemm_bf16x3_kernel<BN>, Cfg::kSmemBytes, cfg, "gemm_bf16x3_kernel")) return 1;
    gemm_bf16x3_kernel<BN><<<grid, NTHREADS, Cfg::kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], p);
    return check_launch("gemm_bf16x3_kernel");
}

int sm_count() {
    static int n[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    dev &= 63;
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

template <int BN>
int launch_pair(const CUtensorMap* maps, const GemmParams& p_in, dim3 grid, cudaStream_t st) {
    using Cfg = PairCfg<BN>;
    static SmemCfg cfg;
    if (ensure_dyn_smem(gemm_pair_kernel<BN>, Cfg::kSmemBytes, cfg, "gemm_pair_kernel")) return 1;
    GemmParams p = p_in;
    p.stats = g_stats;
    gemm_pair_kernel<BN><<<grid, NTHREADS, Cfg::kSmemBytes, st>>>(maps[0], maps[1], maps[2], maps[3], p);
    return check_launch("gemm_pair_kernel");
}

// Width of the CTA-pair tile (0 = use the single-CTA kernel): the pair kernel needs a problem large enough to fill the
// chip with 256-row tiles; among {256, 160} the width with the fewest (waves x width) wins (160 = 640 / 4 = 1280 / 8).
int pick_pair_bn(long n_out, long row_tiles) {
    const char* off = getenv("DFOLD_GEMM_NO_PAIR");
    if (off && off[0] == '1') return 0;
    if (n_out < 160 || row_tiles < 16) return 0;
    const long pairs = cdiv(row_tiles, 2), slots = sm_count() / 2;
    long best = 0, best_cost = 0;
    const int widths[2] = {256, 160};
    for (int w : widths) {
        const long cost = cdiv(cdiv(n_out, w) * pairs, slots) * w;
        if (best == 0 || cost < best_cost) { best = w; best_cost = cost; }
    }
    return (int)best;
}

void default_batching(GemmParams& p) {
    p.bmod = 1; p.a_f_div = 1; p.a_k_bstride = 0; p.b_k_ofs = 0; p.b_k_bstride = 0; p.b_z_bstride = 0;
    p.o_f_div = 1; p.o_col_bstride = 0; p.f_start = 0; p.b_f_add = 0;
    p.zdiv = 1; p.a_f_mul = 1; p.a_z_mul = 0; p.b_f_mul = 1; p.b_z_mul = 0; p.b_n_zmul = 0;
    p.kb_per_split = p.num_kb; p.atomic = 0;
}


// Tile width: every CTA owns one SM (192 KB of shared memory), so the launch runs in ceil(tiles / SMs) waves whose
// duration is proportional to BN.  Pick the width with the smallest waves x BN (ties -> the wider tile, which re-reads
// the A operand less).  `row_tiles` = 128-row tiles x batches / taps in the other grid dimensions.
int pick_bn(long n_out, long row_tiles) {
    if (n_out <= 64) return 64;
    if (n_out <= 128) return 128;
    const long sms = sm_count();
    const long t256 = cdiv(n_out, 256) * row_tiles, t128 = cdiv(n_out, 128) * row_tiles;
    const long c256 = cdiv(t256, sms) * 256, c128 = cdiv(t128, sms) * 128;
    return (c128 < c256) ? 128 : 256;
}

}  // namespace
}  // namespace dfold

using namespace dfold;

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
// Development aid: when `buf` (device, 4 x int64 per CTA of the next launches) is non-null every GEMM launch records, per
// CTA, the cycles of its mainloop and the cycles its MMA-issuing thread / accumulate warps spent waiting on barriers.
extern "C" int dfold_debug_gemm_stats(long long* buf) {
    g_stats = buf;
    return 0;
}

extern "C" int dfold_gemm_bf16x3(
    const uint16_t* a_hi, const uint16_t* a_lo, long F, long F_out, int f_start, long Nr, long K, long lda,
    const uint16_t* b_hi, const uint16_t* b_lo, long n_out, long ldb, int taps_f, int taps_n,
    float* out, long ldo, const float* bias, const float* residual, long ldr,
    float alpha, float beta, int act, void* stream) {
    DFOLD_REQUIRE(F > 0 && F_out > 0 && Nr > 0 && K > 0 && n_out > 0, "dfold_gemm_bf16x3: empty problem");
    DFOLD_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "dfold_gemm_bf16x3: lda/ldb must be multiples of 8 (got %ld, %ld)", lda, ldb);
    DFOLD_REQUIRE(taps_f >= 1 && taps_n >= 1 && (taps_f & 1) && (taps_n & 1), "dfold_gemm_bf16x3: tap grid must be odd");
    const long row_tiles = F_out * cdiv(Nr, BM);
    const int pbn = pick_pair_bn(n_out, row_tiles);
    const int bn = pbn ? pbn : pick_bn(n_out, row_tiles);
    CUtensorMap maps[4];
    // A: dims (K, Nr, F)   B: dims (K, n_out, taps); the pair kernel stages half a B tile per CTA
    if (make_map(&maps[0], a_hi, K, Nr, F, lda, Nr * lda, BK, BM, 1)) return 1;
    if (make_map(&maps[1], a_lo, K, Nr, F, lda, Nr * lda, BK, BM, 1)) return 1;
    const long taps = (long)taps_f * taps_n;
    const int bbox = pbn ? pbn / 2 : bn;
    if (make_map(&maps[2], b_hi, K, n_out, taps, ldb, n_out * ldb, BK, bbox, 1)) return 1;
    if (make_map(&maps[3], b_lo, K, n_out, taps, ldb, n_out * ldb, BK, bbox, 1)) return 1;
    GemmParams p{};
    p.mode = 0;
    p.kc = (int)cdiv(K, BK);
    p.num_kb = (int)(taps * p.kc);
    p.taps_n = taps_n; p.taps_f = taps_f;
    p.tiles_per_frame = (int)cdiv(Nr, BM);
    p.Nr = (int)Nr;
    p.out_rows = F_out * Nr;
    p.n_out = (int)n_out;
    p.out = out; p.ldo = ldo; p.out_tap_stride = 0;
    p.bias = bias; p.res = residual; p.ldr = ldr;
    p.alpha = alpha; p.beta = beta; p.act = act;
    default_batching(p);
    p.f_start = f_start;
    cudaStream_t st = as_stream(stream);
    if (pbn) {
        dim3 pgrid((unsigned)(2 * cdiv(row_tiles, 2)), (unsigned)cdiv(n_out, pbn), 1);
        {
            // activation planes (hi + lo) that fit in L2 next to a weight slice: rows fastest (each column sweep re-reads them
            // from L2); larger than that: columns fastest (measured, ncu dram bytes: 1280->640 conv 772 -> 408 MB per launch,
            // 640->1280 conv 173 MB rows-fastest vs 608 MB columns-fastest)
            const char* e = getenv("DFOLD_GEMM_RASTER");
            const double a_bytes = 4.0 * (double)F * (double)Nr * (double)lda;
            p.raster_n = e ? (e[0] == 'n') : (a_bytes > 64e6);
        }
        if (pbn == 256) return launch_pair<256>(maps, p, pgrid, st);
        return launch_pair<160>(maps, p, pgrid, st);
    }
    // Launches that cannot fill the chip but have a long K loop (the last layers of the dead-frame pyramid: one to a few
    // frames x 25 taps x C_in) are latency-bound on that loop: split K over blockIdx.z, accumulate with atomics into a
    // zeroed output, then apply bias / activation / residual in a second pass.
    const long tiles = cdiv(n_out, bn) * row_tiles;
    int splits = 1;
    {
        const char* off = getenv("DFOLD_GEMM_NO_SPLITK");
        if (!(off && off[0] == '1') && tiles * 2 <= sm_count() && p.num_kb >= 32 && ldo == n_out) {
            splits = (int)min((long)16, min(cdiv((long)sm_count(), tiles), (long)p.num_kb / 8));
            if (splits < 2) splits = 1;
        }
    }
    dim3 grid((unsigned)cdiv(n_out, bn), (unsigned)(F_out * p.tiles_per_frame), (unsigned)splits);
    if (splits > 1) {
        p.kb_per_split = (int)cdiv(p.num_kb, splits);
        grid.z = (unsigned)cdiv(p.num_kb, p.kb_per_split);
        p.atomic = 1;
        p.bias = nullptr; p.res = nullptr; p.act = 0; p.beta = 0.f;
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(F_out * Nr) * (size_t)ldo, st);
        DFOLD_REQUIRE(e == cudaSuccess, "dfold_gemm_bf16x3: memset: %s", cudaGetErrorString(e));
    }
    int rc;
    if (bn == 256) rc = launch<256>(maps, p, grid, st);
    else if (bn == 128) rc = launch<128>(maps, p, grid, st);
    else rc = launch<64>(maps, p, grid, st);
    if (rc || splits == 1) return rc;
    if (bias || residual || act) {
        const long total = F_out * Nr * n_out;
        gemm_finish_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(out, ldo, bias, residual, ldr, beta, act, F_out * Nr, (int)n_out);
        return check_launch("gemm_finish_kernel");
    }
    return 0;
}

// out[tap][m][n] = alpha * sum_{f, j} A[f][j][m] * B[f + df(tap)][j + dn(tap)][n]      (weight gradient; K = pixels)
// A planes [F][Nr][lda] (e.g. the gated output gradient, m = output channel), B planes [Fb][Nr][ldb] (the layer input,
// n = input channel, frame f of A pairs with frame f + b_f_add of B: cropped convolutions): the SAME pixel-major
// planes the forward / data-gradient GEMMs use, read MN-major.
extern "C" int dfold_gemm_wgrad_bf16x3(
    const uint16_t* a_hi, const uint16_t* a_lo, long M, long lda,
    const uint16_t* b_hi, const uint16_t* b_lo, long Nn, long ldb,
    long F, long Fb, int b_f_add, long Nr, int taps_f, int taps_n,
    float* out, long ldo, float alpha, void* stream) {
    DFOLD_REQUIRE(M > 0 && Nn > 0 && F > 0 && Fb > 0 && Nr > 0, "dfold_gemm_wgrad_bf16x3: empty problem");
    DFOLD_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "dfold_gemm_wgrad_bf16x3: lda/ldb must be multiples of 8 (got %ld, %ld)", lda, ldb);
    // ---- CTA-pair path: 256 x {256,128} tiles; the operand whose channel count is an even number of 128-row tiles
    //      takes the M role (swapping roles transposes the store, which the epilogue does with strides) ----
    {
        const char* off = getenv("DFOLD_GEMM_NO_PAIR");
        const bool enabled = !(off && off[0] == '1');
        const long mt = cdiv(M, BM), nt = cdiv(Nn, BM);
        const bool swap = (mt % 2 != 0) && (nt % 2 == 0);
        if (enabled && F * cdiv(Nr, BK) >= 32 && M >= 256 && Nn >= 128 && (mt % 2 == 0 || swap)) {
            const long Mp = swap ? Nn : M, Np = swap ? M : Nn;
            const int pbn = (Np % 256 == 0) ? 256 : 128;
            const uint16_t *pa_hi = swap ? b_hi : a_hi, *pa_lo = swap ? b_lo : a_lo, *pb_hi = swap ? a_hi : b_hi, *pb_lo = swap ? a_lo : b_lo;
            const long pa_ld = swap ? ldb : lda, pb_ld = swap ? lda : ldb, pa_F = swap ? Fb : F, pb_F = swap ? F : Fb;
            CUtensorMap maps[4];
            if (make_map(&maps[0], pa_hi, Mp, Nr, pa_F, pa_ld, Nr * pa_ld, 64, BK, 1)) return 1;
            if (make_map(&maps[1], pa_lo, Mp, Nr, pa_F, pa_ld, Nr * pa_ld, 64, BK, 1)) return 1;
            if (make_map(&maps[2], pb_hi, Np, Nr, pb_F, pb_ld, Nr * pb_ld, 64, BK, 1)) return 1;
            if (make_map(&maps[3], pb_lo, Np, Nr, pb_F, pb_ld, Nr * pb_ld, 64, BK, 1)) return 1;
            GemmParams p{};
            p.mode = 1;
            p.kc = (int)cdiv(Nr, BK);
            p.num_kb = (int)(F * p.kc);
            p.taps_n = taps_n; p.taps_f = taps_f;
            p.tiles_per_frame = 1; p.Nr = (int)Nr;
            p.out_rows = Mp; p.n_out = (int)Np;
            p.out = out; p.ldo = ldo; p.out_tap_stride = M * ldo;
            p.alpha = alpha; p.beta = 0.f; p.act = 0;
            default_batching(p);
            p.b_f_add = b_f_add;
            p.shift_on_a = swap ? 1 : 0;
            p.o_rs = swap ? 1 : ldo;
            p.o_cs = swap ? ldo : 1;
            dim3 pgrid((unsigned)(2 * cdiv(cdiv(Mp, BM), 2)), (unsigned)cdiv(Np, pbn), (unsigned)(taps_f * taps_n));
            cudaStream_t st = as_stream(stream);
            if (pbn == 256) return launch_pair<256>(maps, p, pgrid, st);
            return launch_pair<128>(maps, p, pgrid, st);
        }
    }
    const int bn = pick_bn(Nn, cdiv(M, BM) * taps_f * taps_n);
    CUtensorMap maps[4];
    // dims (channel, residue, frame); boxes of 64 channels x 64 residues
    if (make_map(&maps[0], a_hi, M, Nr, F, lda, Nr * lda, 64, BK, 1)) return 1;
    if (make_map(&maps[1], a_lo, M, Nr, F, lda, Nr * lda, 64, BK, 1)) return 1;
    if (make_map(&maps[2], b_hi, Nn, Nr, Fb, ldb, Nr * ldb, 64, BK, 1)) return 1;
    if (make_map(&maps[3], b_lo, Nn, Nr, Fb, ldb, Nr * ldb, 64, BK, 1)) return 1;
    GemmParams p{};
    p.mode = 1;
    p.kc = (int)cdiv(Nr, BK);
    p.num_kb = (int)(F * p.kc);
    p.taps_n = taps_n; p.taps_f = taps_f;
    p.tiles_per_frame = 1; p.Nr = (int)Nr;
    p.out_rows = M; p.n_out = (int)Nn;
    p.out = out; p.ldo = ldo; p.out_tap_stride = M * ldo;
    p.bias = nullptr; p.res = nullptr; p.ldr = 0;
    p.alpha = alpha; p.beta = 0.f; p.act = 0;
    default_batching(p);
    p.b_f_add = b_f_add;
    dim3 grid((unsigned)cdiv(Nn, bn), (unsigned)cdiv(M, BM), (unsigned)(taps_f * taps_n));
    cudaStream_t st = as_stream(stream);
    if (bn == 256) return launch<256>(maps, p, grid, st);
    if (bn == 128) return launch<128>(maps, p, grid, st);
    return launch<64>(maps, p, grid, st);
}

// Batched K-major GEMM (no taps): for every row tile of "frame" b in [0, n_batches), hb = b % bmod:
//   out[(b / o_f_div) * Nr + n, hb * o_col_bstride + c] = alpha * sum_k A[b / a_f_div][n][hb * a_k_bstride + k]
//                                                                    * B[hb * b_z_bstride][c][b_k_ofs + hb * b_k_bstride + k]
// A planes: [a_frames][Nr][lda]; B planes: [b_z][b_rows][ldb].
extern "C" int dfold_gemm_bf16x3_batched(
    const uint16_t* a_hi, const uint16_t* a_lo, long a_frames, long Nr, long a_cols, long lda,
    long n_batches, int bmod, int a_f_div, long a_k_bstride, long K,
    const uint16_t* b_hi, const uint16_t* b_lo, long b_z, long b_rows, long b_cols, long ldb,
    long b_k_ofs, long b_k_bstride, int b_z_bstride, long n_out,
    float* out, long ldo, long out_rows, int o_f_div, long o_col_bstride, float alpha, void* stream) {
    DFOLD_REQUIRE(n_batches > 0 && Nr > 0 && K > 0 && n_out > 0 && bmod >= 1, "dfold_gemm_bf16x3_batched: empty problem");
    DFOLD_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "dfold_gemm_bf16x3_batched: lda/ldb must be multiples of 8");
    DFOLD_REQUIRE(a_k_bstride % 8 == 0 && b_k_ofs % 8 == 0 && b_k_bstride % 8 == 0,
                  "dfold_gemm_bf16x3_batched: K offsets must be multiples of 8 elements (TMA 16-byte alignment)");
    const int bn = pick_bn(n_out, n_batches * cdiv(Nr, BM));
    CUtensorMap maps[4];
    if (make_map(&maps[0], a_hi, a_cols, Nr, a_frames, lda, Nr * lda, BK, BM, 1)) return 1;
    if (make_map(&maps[1], a_lo, a_cols, Nr, a_frames, lda, Nr * lda, BK, BM, 1)) return 1;
    if (make_map(&maps[2], b_hi, b_cols, b_rows, b_z, ldb, b_rows * ldb, BK, bn, 1)) return 1;
    if (make_map(&maps[3], b_lo, b_cols, b_rows, b_z, ldb, b_rows * ldb, BK, bn, 1)) return 1;
    GemmParams p{};
    p.mode = 0;
    p.kc = (int)cdiv(K, BK);
    p.num_kb = p.kc;
    p.taps_n = 1; p.taps_f = 1;
    p.tiles_per_frame = (int)cdiv(Nr, BM);
    p.Nr = (int)Nr;
    p.out_rows = out_rows;
    p.n_out = (int)n_out;
    p.out = out; p.ldo = ldo; p.out_tap_stride = 0;
    p.bias = nullptr; p.res = nullptr; p.ldr = 0;
    p.alpha = alpha; p.beta = 0.f; p.act = 0;
    default_batching(p);
    p.bmod = bmod; p.a_f_div = a_f_div; p.a_k_bstride = (int)a_k_bstride;
    p.b_k_ofs = (int)b_k_ofs; p.b_k_bstride = (int)b_k_bstride; p.b_z_bstride = b_z_bstride;
    p.o_f_div = o_f_div; p.o_col_bstride = o_col_bstride;
    // K tail: the A / B boxes may run past [k_ofs, k_ofs + K) into the neighbouring batch slice; require K % 64 == 0
    DFOLD_REQUIRE(K % BK == 0 || (a_k_bstride == 0 && b_k_ofs == 0 && b_k_bstride == 0),
                  "dfold_gemm_bf16x3_batched: K must be a multiple of 64 when it is a slice of a wider row");
    dim3 grid((unsigned)cdiv(n_out, bn), (unsigned)(n_batches * p.tiles_per_frame), 1);
    cudaStream_t st = as_stream(stream);
    if (bn == 256) return launch<256>(maps, p, grid, st);
    if (bn == 128) return launch<128>(maps, p, grid, st);
    return launch<64>(maps, p, grid, st);
}

// Batched / split-K MN-major GEMM (K = rows):  for z = zq * splits + zs
//   out[zq][m][n] (+)= alpha * sum_{f in split zs} sum_j A[f * a_f_mul + zq * a_z_mul][j][m] * B[f * b_f_mul + zq * b_z_mul][j][zq * b_n_zmul + n]
// A planes: dims (M, a_mid, a_outer) with element strides (lda, a_ostride); B likewise.  K loop: Fk outer x ceil(Nr/64) blocks.
extern "C" int dfold_gemm_wgrad_bf16x3_batched(
    const uint16_t* a_hi, const uint16_t* a_lo, long M, long a_mid, long a_outer, long lda, long a_ostride,
    const uint16_t* b_hi, const uint16_t* b_lo, long b_cols, long b_mid, long b_outer, long ldb, long b_ostride,
    long Nn, long Fk, long Nr, int zcount, int splits, int a_f_mul, int a_z_mul, int b_f_mul, int b_z_mul, long b_n_zmul,
    float* out, long ldo, long out_z_stride, float alpha, void* stream) {
    DFOLD_REQUIRE(M > 0 && Nn > 0 && Fk > 0 && Nr > 0 && zcount > 0 && splits > 0, "dfold_gemm_wgrad_bf16x3_batched: empty problem");
    DFOLD_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && a_ostride % 8 == 0 && b_ostride % 8 == 0 && b_n_zmul % 8 == 0,
                  "dfold_gemm_wgrad_bf16x3_batched: strides / offsets must be multiples of 8 elements");
    const int bn = pick_bn(Nn, cdiv(M, BM) * zcount * splits);
    CUtensorMap maps[4];
    if (make_map(&maps[0], a_hi, M, a_mid, a_outer, lda, a_ostride, 64, BK, 1)) return 1;
    if (make_map(&maps[1], a_lo, M, a_mid, a_outer, lda, a_ostride, 64, BK, 1)) return 1;
    if (make_map(&maps[2], b_hi, b_cols, b_mid, b_outer, ldb, b_ostride, 64, BK, 1)) return 1;
    if (make_map(&maps[3], b_lo, b_cols, b_mid, b_outer, ldb, b_ostride, 64, BK, 1)) return 1;
    GemmParams p{};
    p.mode = 1;
    p.kc = (int)cdiv(Nr, BK);
    p.num_kb = (int)(Fk * p.kc);
    p.taps_n = 1; p.taps_f = 1;
    p.tiles_per_frame = 1; p.Nr = (int)Nr;
    p.out_rows = M; p.n_out = (int)Nn;
    p.out = out; p.ldo = ldo; p.out_tap_stride = out_z_stride;
    p.bias = nullptr; p.res = nullptr; p.ldr = 0;
    p.alpha = alpha; p.beta = 0.f; p.act = 0;
    default_batching(p);
    p.zdiv = splits;
    p.a_f_mul = a_f_mul; p.a_z_mul = a_z_mul; p.b_f_mul = b_f_mul; p.b_z_mul = b_z_mul; p.b_n_zmul = (int)b_n_zmul;
    p.kb_per_split = (int)cdiv(p.num_kb, splits);
    p.atomic = splits > 1;
    DFOLD_REQUIRE((long)zcount * splits <= 65535, "dfold_gemm_wgrad_bf16x3_batched: too many z blocks");
    dim3 grid((unsigned)cdiv(Nn, bn), (unsigned)cdiv(M, BM), (unsigned)(zcount * splits));
    cudaStream_t st = as_stream(stream);
    if (bn == 256) return launch<256>(maps, p, grid, st);
    if (bn == 128) return launch<128>(maps, p, grid, st);
    return launch<64>(maps, p, grid, st);
}
