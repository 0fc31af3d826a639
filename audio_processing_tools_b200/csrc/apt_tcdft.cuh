// apt_tcdft.cuh -- the STFT of the 256/128 geometry as a DFT-by-GEMM on the 5th-generation tensor cores (tcgen05):
// the "DFT-as-GEMM tensor-core variant" of north_star subsystem 2 (reference site: edge/rain_signal_processor.py:818-826).
//
//   S[t, k] = sum_n x[128 (t - 1) + n] * w[n] * exp(-2 pi i k n / 256),   k in the operating band only
//
// is one GEMM  D[frames x 2K] = A[frames x 256] * B[256 x 2K]  (B = window x twiddles: the real and imaginary column of every
// band bin; 2K = 142 -> N = 144).  Arithmetic: the int16 sample is split exactly into two 8-bit limbs, x = 256 xh + xl, both
// exact in fp16; B is split into two fp16 limbs (B ~ B1 + B2, |error| < 2^-22 of the largest coefficient); the products run
// as kind::f16 MMAs with fp32 accumulation in tensor memory, one accumulator per sample limb (D_hi, D_lo), combined in the
// epilogue:  S = (256 D_hi + D_lo) / (1024 * 32767).  Accuracy class of the float32 FFT (spectra ~1e-6 of the frame maximum):
// this is the tolerance path (fft = "tc"), not the bit-exact float64 default.
//
// One CTA per SM, persistent over tiles of 128 frames:
//   B (144 KB, both limbs, all of K) is brought in ONCE per CTA by bulk-tensor-style TMA copies (cp.async.bulk, mbarrier
//   complete_tx) into the canonical K-major SWIZZLE_128B layout the MMA reads (the image is prepared on the host);
//   A is produced per tile in four K-chunks of 64 samples by the 128 worker threads: coalesced 128-bit loads of the raw int16
//   samples (overlapping frames come from L1/L2), limb split with two logic + two half2 operations per sample pair, 128-bit
//   stores into the same swizzled layout, double buffered against the MMA through full/empty mbarriers;
//   one elected thread issues the tcgen05.mma instructions (M = 128, N = 144, K = 16 each: 4 per K-step) and commits to the
//   mbarriers; the 128 worker threads then read the accumulators with tcgen05.ld (thread = frame row), form |S|^2 and write
//   the band plane (through shared memory, coalesced) and / or the band energies.
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
constexpr int TC_SMEM = TC_B_BYTES + TC_A_BYTES + 1024 /* barriers + tmem pointer */ + 1024 /* alignment slack */;
constexpr int TC_NT = 160;           // 4 worker warps + 1 MMA / TMA warp
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
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + TC_A_BYTES);
    uint64_t* bar_b = bars + 0;          // B image landed
    uint64_t* bar_full = bars + 1;       // [2] A slot filled (128 arrivals)
    uint64_t* bar_empty = bars + 3;      // [2] A slot consumed by the MMAs (tcgen05.commit)
    uint64_t* bar_acc = bars + 5;        // accumulators complete (tcgen05.commit)
    uint64_t* bar_accfree = bars + 6;    // accumulators read out (128 arrivals)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = warp < 4;

    if (tid == 0) {
        tc_mbar_init(bar_b, 1);
        tc_mbar_init(bar_full + 0, 128); tc_mbar_init(bar_full + 1, 128);
        tc_mbar_init(bar_empty + 0, 1); tc_mbar_init(bar_empty + 1, 1);
        tc_mbar_init(bar_acc, 1);
        tc_mbar_init(bar_accfree, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    const uint32_t tm_hi = tmem, tm_lo = tmem + 256;         // accumulator columns [0,144) and [256,400)

    // the B image: eight TMA bulk copies (limb x chunk), one mbarrier
    if (tid == 128) {
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
    uint32_t it_fill = 0, it_mma = 0;        // chunks filled / consumed so far (slot = it & 1, phase = (it >> 1) & 1)
    uint32_t n_tile_done = 0;

    for (int tile = tile_lo + (int)blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, n_tile_done++) {
        const int t0 = tile * TC_M;
        if (worker) {
            // ---- A producer: chunk kc = samples 64 kc .. 64 kc + 63 of every frame of the tile
            for (int kc = 0; kc < TC_NCHUNK && ok; kc++, it_fill++) {
                const int slot = it_fill & 1;
                if (it_fill >= 2) ok = tc_mbar_wait(bar_empty + slot, ((it_fill >> 1) - 1) & 1);
                if (!ok) break;
                unsigned char* a_hi = sA + (size_t)(slot * 2 + 0) * TC_A_TILE;
                unsigned char* a_lo = sA + (size_t)(slot * 2 + 1) * TC_A_TILE;
#pragma unroll 2
                for (int item = tid; item < TC_M * 8; item += 128) {
                    const int r = item >> 3, c16 = item & 7;
                    const int t = t0 + r;
                    const int64_t s = (int64_t)(t - 1) * 128 + kc * TC_KC + c16 * 8;     // first of 8 samples (clip-relative)
                    uint32_t w[4] = {0u, 0u, 0u, 0u};                                      // 8 int16, little endian pairs
                    if (t < T_clip) {
                        if constexpr (sizeof(PCM) == 2) {
                            if (aligned && s >= 0 && s + 8 <= N) {
                                const uint4 v = __ldg(reinterpret_cast<const uint4*>(pcm + base + s));
                                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; e++) {
                                    const int64_t se = s + e;
                                    const uint32_t h = (se >= 0 && se < N) ? (uint32_t)(uint16_t)__ldg(pcm + base + se) : 0u;
                                    w[e >> 1] |= h << (16 * (e & 1));
                                }
                            }
                        }
                    }
                    // limb split: xl = low byte (0..255), xh = high byte as signed (-128..127); 0x6400 | n is the fp16 1024 + n
                    uint32_t lo[4], hi[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t l = (w[i] & 0x00ff00ffu) | 0x64006400u;
                        const uint32_t h = (((w[i] >> 8) & 0x00ff00ffu) ^ 0x00800080u) | 0x64006400u;
                        const __half2 lh = __hsub2(*reinterpret_cast<const __half2*>(&l), __floats2half2_rn(1024.0f, 1024.0f));
                        const __half2 hh = __hsub2(*reinterpret_cast<const __half2*>(&h), __floats2half2_rn(1152.0f, 1152.0f));
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
        } else if (tid == 128) {
            // ---- MMA issuer
            if (n_tile_done == 0) ok = tc_mbar_wait(bar_b, 0);
            if (ok && n_tile_done > 0) ok = tc_mbar_wait(bar_accfree, (n_tile_done - 1) & 1);     // epilogue of the previous tile done
            for (int kc = 0; kc < TC_NCHUNK && ok; kc++, it_mma++) {
                const int slot = it_mma & 1;
                ok = tc_mbar_wait(bar_full + slot, (it_mma >> 1) & 1);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = tc_smem_u32(sA + (size_t)(slot * 2 + 0) * TC_A_TILE);
                const uint32_t a_lo = tc_smem_u32(sA + (size_t)(slot * 2 + 1) * TC_A_TILE);
                const uint32_t b1 = tc_smem_u32(sB + (size_t)(0 * TC_NCHUNK + kc) * TC_B_TILE);
                const uint32_t b2 = tc_smem_u32(sB + (size_t)(1 * TC_NCHUNK + kc) * TC_B_TILE);
#pragma unroll
                for (int ks = 0; ks < TC_KC / 16; ks++) {
                    const uint32_t ko = ks * 32;                      // 16 fp16 = 32 bytes along K inside the swizzled row
                    const uint32_t acc = (kc | ks) ? 1u : 0u;
                    tc_mma(tm_hi, tc_desc(a_hi + ko), tc_desc(b1 + ko), acc);
                    tc_mma(tm_hi, tc_desc(a_hi + ko), tc_desc(b2 + ko), 1u);
                    tc_mma(tm_lo, tc_desc(a_lo + ko), tc_desc(b1 + ko), acc);
                    tc_mma(tm_lo, tc_desc(a_lo + ko), tc_desc(b2 + ko), 1u);
                }
                tc_commit(bar_empty + slot);                          // slot free once these MMAs have read it
            }
            if (ok) tc_commit(bar_acc);
        }
        if (worker && ok) {
            // ---- epilogue: thread = frame row; TMEM lane = 32 * warp + lane
            ok = tc_mbar_wait(bar_acc, n_tile_done & 1);
            if (ok) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int r = tid;
                const int t = t0 + r;
                const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
                const float k1 = 256.0f, k2 = 1.0f / (TC_BSCALE * 32767.0f);
                float be[APT_MAX_MODES + 1];
#pragma unroll
                for (int m = 0; m <= APT_MAX_MODES; m++) be[m] = 0.0f;
                // the band plane goes through the (now idle) A slots: [row][K] floats, then a coalesced copy out
                float* sP = reinterpret_cast<float*>(sA);
                const int K = q.K;
                for (int cb = 0; cb < TC_N; cb += 16) {
                    float vh[16], vl[16];
                    tc_ld16(tm_hi + lane_off + cb, vh);
                    tc_ld16(tm_lo + lane_off + cb, vl);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int kb = (cb >> 1) + j;
                        if (kb < K) {
                            const float re = fmaf(k1, vh[2 * j], vl[2 * j]) * k2;
                            const float im = fmaf(k1, vh[2 * j + 1], vl[2 * j + 1]) * k2;
                            const float pw = fmaf(re, re, im * im);
                            if (q.P_band) sP[r * K + kb] = pw;
                            if (q.band_energy) {
                                be[q.n_modes] += pw;
#pragma unroll
                                for (int m = 0; m < APT_MAX_MODES; m++)
                                    if (m < q.n_modes && kb >= q.mode_blo[m] && kb <= q.mode_bhi[m]) be[m] += pw;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tc_mbar_arrive(bar_accfree);                          // the next tile's MMAs may overwrite the accumulators
                if (q.band_energy && t < T_clip) {
                    for (int m = 0; m < q.n_modes; m++) q.band_energy[(int64_t)m * q.nF + f0 + t] = be[m];
                    q.band_energy[(int64_t)q.n_modes * q.nF + f0 + t] = be[q.n_modes] + q.eps;
                }
                if (q.P_band) {
                    asm volatile("bar.sync 1, 128;" ::: "memory");    // the four worker warps only
                    const int nrow = min(TC_M, T_clip - t0);
                    float* dst = q.P_band + (f0 + t0) * K;
                    for (int i = tid; i < nrow * K; i += 128) dst[i] = sP[i];
                    asm volatile("bar.sync 1, 128;" ::: "memory");    // sP is the next tile's A slot
                }
            }
        }
    }
    if (!ok && q.error_flag) *q.error_flag = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace apt
