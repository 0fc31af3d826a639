// apt_tcdft.cuh -- the STFT of the 256/128 geometry as a DFT-by-GEMM on the 5th-generation tensor cores (tcgen05):
// the "DFT-as-GEMM tensor-core variant" of north_star subsystem 2 (reference site: edge/rain_signal_processor.py:818-826).
//
//   S[t, k] = sum_n x[128 (t - 1) + n] * w[n] * exp(-2 pi i k n / 256),   k in the operating band only
//
// is one GEMM  D[frames x 2K] = A[frames x 256] * B[256 x 2K]  (B = window x twiddles: the real and imaginary column of every
// band bin; 2K = 142 -> N = 144).  Arithmetic: the int16 sample is split exactly into two 8-bit limbs, x = 256 xh + xl, both
// exact in fp16; B is split into two fp16 limbs (B ~ B1 + B2, |error| < 2^-22 of the largest coefficient); the products run
// as kind::f16 MMAs with fp32 accumulation in tensor memory (the high limb enters as 256 xh, so one accumulator takes all
// four products):  S = D / (1024 * 32767).  Accuracy class of the float32 FFT (spectra ~1e-6 of the frame maximum):
// this is the tolerance path (fft = "tc"), not the bit-exact float64 default.
//
// One CTA per SM, persistent over tiles of 128 frames:
//   B (144 KB, both limbs, all of K) is brought in ONCE per CTA by bulk-tensor-style TMA copies (cp.async.bulk, mbarrier
//   complete_tx) into the canonical K-major SWIZZLE_128B layout the MMA reads (the image is prepared on the host);
//   A is produced per tile in four K-chunks of 64 samples by 8 producer warps: coalesced 128-bit loads of the raw int16
//   samples (overlapping frames come from L1/L2), limb split with two logic + two / three half2 operations per sample pair
//   (the high limb is stored as 256 xh, exact in fp16, so that both limbs accumulate into ONE accumulator), 128-bit stores
//   into the same swizzled layout, double buffered against the MMA through full / empty mbarriers;
//   one elected thread issues the tcgen05.mma instructions (M = 128, N = 144, K = 16 each: 4 per K-step) and commits to the
//   mbarriers; the accumulator (144 of 512 TMEM columns) is double buffered, so 4 epilogue warps read tile n with tcgen05.ld
//   (thread = frame row), form |S|^2 and write the band plane / the band energies while the MMAs of tile n + 1 run.
// Every mbarrier wait is bounded: a barrier that never completes sets an error flag and returns instead of hanging the GPU.
#pragma once
#include <cuda_fp16.h>
#include "apt_kernels.cuh"

namespace apt {

constexpr int TC_M = 128;            // frames per tile
constexpr int TC_N = 144;            // 2 * band bins, padded to a multiple of 16
constexpr int TC_KC = 64;            // samples per K-chunk (one 128-byte swizzle row of fp16)
constexpr int TC_NCHUNK = 256 / TC_KC;
constexpr int TC_A_TILE = TC_M * TC_KC * 2;        // 16 KB: one limb of one chunk
constexpr int TC_B_TILE = TC_N * TC_KC * 2;        // 18 KB
constexpr int TC_B_BYTES = 2 * TC_NCHUNK * TC_B_TILE;   // 144 KB
constexpr int TC_A_BYTES = 2 * 2 * TC_A_TILE;      // 2 slots x 2 limbs
constexpr int TC_E_BYTES = 4 * 32 * 9 * 4;         // epilogue staging: per warp 32 rows x (8 + 1) floats
constexpr int TC_SMEM = TC_B_BYTES + TC_A_BYTES + TC_E_BYTES + 1024 /* barriers + tmem pointer */ + 1024 /* alignment slack */;
constexpr int TC_NPROD = 256;        // producer threads (warps 0..7)
constexpr int TC_NT = TC_NPROD + 128 + 32;   // + 4 epilogue warps (8..11) + the MMA / TMA warp (12)
constexpr float TC_BSCALE = 1024.0f; // B is stored times 1024 (keeps the low limb out of fp16's subnormal range)
constexpr uint32_t TC_SPIN_LIMIT = 1u << 26;

struct TcParams {
    const unsigned char* Bimg;   // [TC_B_BYTES] device: B in the shared-memory image (limb, chunk, swizzled rows)
    float* P_band;               // [nF][K] optional
    float* band_energy;          // [M + 1][nF] optional
    int64_t nF;
    int K, band_lo, n_modes;
    int mode_blo[APT_MAX_MODES], mode_bhi[APT_MAX_MODES];   // band-relative mode bins
    float eps;
    int* error_flag;             // set to 1 if a barrier wait ran into its bound
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait on a phase parity; false = gave up
__device__ __forceinline__ bool tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = tc_smem_u32(bar);
    for (uint32_t it = 0; it < TC_SPIN_LIMIT; it++) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc_smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (1 for swizzled K-major) in [16,30), stride byte offset (8 rows x 128 B = 1024 -> 64) in [32,46),
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = F16 (0), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// byte offset of the 16-byte chunk c16 of row r inside a K-major SWIZZLE_128B tile (rows of 128 bytes)
__host__ __device__ __forceinline__ int tc_swz(int r, int c16) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4); }

template <typename PCM>
__global__ void __launch_bounds__(TC_NT, 1) tcdft256_kernel(Batch b, const PCM* __restrict__ pcm, const int64_t* __restrict__ tile_off,
                                                            const TcParams q) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sB = sm;                                  // [limb][chunk][TC_B_TILE]
    unsigned char* sA = sm + TC_B_BYTES;                     // [slot][limb][TC_A_TILE]
    float* sE = reinterpret_cast<float*>(sA + TC_A_BYTES);   // [4 warps][32][9] epilogue staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + TC_A_BYTES + TC_E_BYTES);
    uint64_t* bar_b = bars + 0;          // B image landed
    uint64_t* bar_full = bars + 1;       // [2] A slot filled (TC_NPROD arrivals)
    uint64_t* bar_empty = bars + 3;      // [2] A slot consumed by the MMAs (tcgen05.commit)
    uint64_t* bar_acc = bars + 5;        // [2] accumulator stage complete (tcgen05.commit)
    uint64_t* bar_accfree = bars + 7;    // [2] accumulator stage read out (128 arrivals)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 10);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = warp < 8, epilogue = warp >= 8 && warp < 12, mma_thread = tid == TC_NPROD + 128;

    if (tid == 0) {
        tc_mbar_init(bar_b, 1);
        for (int i = 0; i < 2; i++) {
            tc_mbar_init(bar_full + i, TC_NPROD); tc_mbar_init(bar_empty + i, 1);
            tc_mbar_init(bar_acc + i, 1); tc_mbar_init(bar_accfree + i, 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;                           // accumulator stage s: columns [256 s, 256 s + 144)

    // the B image: eight TMA bulk copies (limb x chunk), one mbarrier
    if (mma_thread) {
        tc_mbar_expect_tx(bar_b, (uint32_t)TC_B_BYTES);
        for (int i = 0; i < 2 * TC_NCHUNK; i++) tc_bulk_g2s(sB + (size_t)i * TC_B_TILE, q.Bimg + (size_t)i * TC_B_TILE, TC_B_TILE, bar_b);
    }

    const int c = b.clip0 + (int)blockIdx.y;
    const int64_t base = __ldg(b.samp_off + c);
    const int64_t N = __ldg(b.samp_off + c + 1) - base;
    const int64_t f0 = __ldg(b.frame_off + c);
    const int T_clip = (int)(__ldg(b.frame_off + c + 1) - f0);
    // tiles of 128 frames inside the launch's frame range [b.ta, b.tb) (time segments are multiples of 128 frames)
    const int tile_lo = b.ta / TC_M;
    const int n_tiles = (int)((min((int64_t)T_clip, (int64_t)b.tb) + TC_M - 1) / TC_M);
    const bool aligned = ((base & 7) == 0) && sizeof(PCM) == 2;
    bool ok = true;
    uint32_t it = 0;                 // K-chunks filled (producers) / consumed (MMA thread) so far: slot = it & 1
    uint32_t nt = 0;                 // tiles done: accumulator stage = nt & 1

    if (producer) {
        // ---- A producers: chunk kc = samples 64 kc .. 64 kc + 63 of every frame of the tile, both limbs
        for (int tile = tile_lo + (int)blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
            const int t0 = tile * TC_M;
            for (int kc = 0; kc < TC_NCHUNK && ok; kc++, it++) {
                const int slot = it & 1;
                // all four loads of this thread's items first (in flight while the slot is still being read)
                uint4 raw[4];
                bool fast[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int item = tid + u * TC_NPROD;
                    const int r = item >> 3, c16 = item & 7;
                    const int64_t s = (int64_t)(t0 + r - 1) * 128 + kc * TC_KC + c16 * 8;     // first of 8 samples (clip-relative)
                    fast[u] = aligned && (t0 + r) < T_clip && s >= 0 && s + 8 <= N;
                    raw[u] = make_uint4(0u, 0u, 0u, 0u);
                    if constexpr (sizeof(PCM) == 2) {
                        if (fast[u]) raw[u] = __ldg(reinterpret_cast<const uint4*>(pcm + base + s));
                        else if ((t0 + r) < T_clip) {
                            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                            for (int e = 0; e < 8; e++) {
                                const int64_t se = s + e;
                                const uint32_t h = (se >= 0 && se < N) ? (uint32_t)(uint16_t)__ldg(pcm + base + se) : 0u;
                                w[e >> 1] |= h << (16 * (e & 1));
                            }
                            raw[u] = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
                if (it >= 2) ok = tc_mbar_wait(bar_empty + slot, ((it >> 1) - 1) & 1);
                if (!ok) break;
                unsigned char* a_hi = sA + (size_t)(slot * 2 + 0) * TC_A_TILE;
                unsigned char* a_lo = sA + (size_t)(slot * 2 + 1) * TC_A_TILE;
                const __half2 c1024 = __floats2half2_rn(1024.0f, 1024.0f), c1152 = __floats2half2_rn(1152.0f, 1152.0f);
                const __half2 c256 = __floats2half2_rn(256.0f, 256.0f);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int item = tid + u * TC_NPROD;
                    const int r = item >> 3, c16 = item & 7;
                    const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
                    // limb split: xl = low byte (0..255), xh = high byte as signed (-128..127); 0x6400 | n is the fp16 1024 + n.
                    // The high limb is stored as 256 * xh (exact in fp16), so that both limbs accumulate into one accumulator.
                    uint32_t lo[4], hi[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t l = (w[i] & 0x00ff00ffu) | 0x64006400u;
                        const uint32_t h = (((w[i] >> 8) & 0x00ff00ffu) ^ 0x00800080u) | 0x64006400u;
                        const __half2 lh = __hsub2(*reinterpret_cast<const __half2*>(&l), c1024);
                        const __half2 hh = __hmul2(__hsub2(*reinterpret_cast<const __half2*>(&h), c1152), c256);
                        lo[i] = *reinterpret_cast<const uint32_t*>(&lh);
                        hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    const int off = tc_swz(r, c16);
                    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA
                tc_mbar_arrive(bar_full + slot);
            }
        }
    } else if (mma_thread) {
        // ---- MMA issuer: per K-chunk 4 K-steps x (hi, lo) x (B1, B2) into the tile's accumulator stage
        ok = tc_mbar_wait(bar_b, 0);
        for (int tile = tile_lo + (int)blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, nt++) {
            const uint32_t stage = nt & 1;
            if (nt >= 2) ok = tc_mbar_wait(bar_accfree + stage, ((nt >> 1) - 1) & 1);     // epilogue of tile nt - 2 has drained this stage
            const uint32_t tm = tmem + 256u * stage;
            for (int kc = 0; kc < TC_NCHUNK && ok; kc++, it++) {
                const int slot = it & 1;
                ok = tc_mbar_wait(bar_full + slot, (it >> 1) & 1);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = tc_smem_u32(sA + (size_t)(slot * 2 + 0) * TC_A_TILE);
                const uint32_t a_lo = tc_smem_u32(sA + (size_t)(slot * 2 + 1) * TC_A_TILE);
                const uint32_t b1 = tc_smem_u32(sB + (size_t)(0 * TC_NCHUNK + kc) * TC_B_TILE);
                const uint32_t b2 = tc_smem_u32(sB + (size_t)(1 * TC_NCHUNK + kc) * TC_B_TILE);
#pragma unroll
                for (int ks = 0; ks < TC_KC / 16; ks++) {
                    const uint32_t ko = ks * 32;                      // 16 fp16 = 32 bytes along K inside the swizzled row
                    tc_mma(tm, tc_desc(a_hi + ko), tc_desc(b1 + ko), (kc | ks) ? 1u : 0u);
                    tc_mma(tm, tc_desc(a_hi + ko), tc_desc(b2 + ko), 1u);
                    tc_mma(tm, tc_desc(a_lo + ko), tc_desc(b1 + ko), 1u);
                    tc_mma(tm, tc_desc(a_lo + ko), tc_desc(b2 + ko), 1u);
                }
                tc_commit(bar_empty + slot);                          // slot free once these MMAs have read it
            }
            if (ok) tc_commit(bar_acc + stage);
        }
    } else if (epilogue) {
        // ---- epilogue: thread = frame row; TMEM lane = 32 * (warp - 8) + lane
        const int ew = warp - 8, r = ew * 32 + lane;
        const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
        float* st = sE + ew * 32 * 9;
        const float k2 = 1.0f / (TC_BSCALE * 32767.0f);
        const int K = q.K;
        for (int tile = tile_lo + (int)blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, nt++) {
            const int t0 = tile * TC_M;
            const int t = t0 + r;
            const uint32_t stage = nt & 1;
            ok = tc_mbar_wait(bar_acc + stage, (nt >> 1) & 1);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tm = tmem + 256u * stage + lane_off;
            float be[APT_MAX_MODES + 1];
#pragma unroll
            for (int m = 0; m <= APT_MAX_MODES; m++) be[m] = 0.0f;
            const int nrow = min(32, T_clip - (t0 + ew * 32));         // rows of this warp that exist
            for (int cb = 0; cb < TC_N; cb += 16) {
                float v[16];
                tc_ld16(tm + cb, v);
                float pw[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float re = v[2 * j] * k2, im = v[2 * j + 1] * k2;
                    pw[j] = fmaf(re, re, im * im);
                }
                const int kb0 = cb >> 1;
                if (q.band_energy) {
                    // the eight bins of this step against every mode band that intersects them (a band spans one or two steps)
                    const int nv = min(8, K - kb0);
#pragma unroll
                    for (int j = 0; j < 8; j++) if (j < nv) be[APT_MAX_MODES] += pw[j];
#pragma unroll
                    for (int m = 0; m < APT_MAX_MODES; m++) {
                        const int lo = max(q.mode_blo[m] - kb0, 0), hi = min(min(q.mode_bhi[m] - kb0, 7), nv - 1);
                        if (m < q.n_modes && lo <= hi) {
#pragma unroll
                            for (int j = 0; j < 8; j++) if (j >= lo && j <= hi) be[m] += pw[j];
                        }
                    }
                }
                if (q.P_band) {
                    // 8 bins of 32 rows through the warp's staging tile, so that the stores cover whole 32-byte runs of a row
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; j++) st[lane * 9 + j] = pw[j];
                    __syncwarp();
                    float* dst = q.P_band + (f0 + t0 + ew * 32) * K + kb0;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int idx = e * 32 + lane, rr = idx >> 3, cc = idx & 7;
                        if (rr < nrow && kb0 + cc < K) dst[(int64_t)rr * K + cc] = st[rr * 9 + cc];
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_mbar_arrive(bar_accfree + stage);                      // the MMAs of tile nt + 2 may overwrite this stage
            if (q.band_energy && t < T_clip) {
                for (int m = 0; m < q.n_modes; m++) q.band_energy[(int64_t)m * q.nF + f0 + t] = be[m];
                q.band_energy[(int64_t)q.n_modes * q.nF + f0 + t] = be[APT_MAX_MODES] + q.eps;
            }
        }
    }
    if (!ok && q.error_flag) *q.error_flag = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace apt
