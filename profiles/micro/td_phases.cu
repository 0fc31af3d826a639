// Phase timing of td_features_kernel on one interior tile (profiling build, -DAPT_PROFILE_PHASES).
#define APT_PROFILE_PHASES 1
#include "../../audio_processing_tools_b200/csrc/apt_b200.cu"
#include <cstdio>
#include <vector>
#include <cmath>
int main() {
    apt_ctx* ctx; if (apt_init(0, &ctx)) { printf("no gpu\n"); return 1; }
    apt_params_t p; apt_params_default(&p);
    p.n_modes = 5; int lo[5] = {11, 19, 35, 54, 73}, hi[5] = {14, 24, 41, 58, 76};
    for (int i = 0; i < 5; i++) { p.mode_lo[i] = lo[i]; p.mode_hi[i] = hi[i]; p.mode_band_lo[i] = lo[i] - 10; p.mode_band_hi[i] = hi[i] - 10; }
    std::vector<double> win(256); std::vector<float> fr(129);
    for (int i = 0; i < 256; i++) win[i] = 0.5 - 0.5 * cos(2 * M_PI * i / 256.0);
    for (int i = 0; i < 129; i++) fr[i] = i * 11162.0f / 256;
    p.window = win.data(); p.freqs = fr.data();
    p.n_sos = 2; p.padlen = 15;
    double sos[2][6] = {{0.7726678130550604, -1.5453356261101208, 0.7726678130550604, 1, -1.6609362580597045, 0.6937015023502847},
                        {1, -2, 1, 1, -1.8246287238899792, 0.8606231249922253}};
    memcpy(p.sos, sos, sizeof(sos)); p.zi[0][0] = -0.77; p.zi[0][1] = 0.77; p.zi[1][0] = 0; p.zi[1][1] = 0;
    const int n_clips = 64; std::vector<int64_t> len(n_clips, 11162 * 60);
    apt_plan_t* pl; if (apt_plan_create(ctx, &p, n_clips, len.data(), &pl)) { printf("plan: %s\n", apt_last_error(ctx)); return 1; }
    int16_t* pcm; cudaMalloc(&pcm, pl->nS * 2);
    std::vector<int16_t> h(pl->nS); for (size_t i = 0; i < h.size(); i++) h[i] = (int16_t)((i * 2654435761u) >> 18) / 8;
    cudaMemcpy(pcm, h.data(), pl->nS * 2, cudaMemcpyHostToDevice);
    long long* dbg; cudaMalloc(&dbg, 16 * 8); cudaMemset(dbg, 0, 128);
    Batch b{0, n_clips, pl->d_samp_off.p, pl->d_frame_off.p};
    TdOut to; to.td = pl->d_td.p; to.x_td = nullptr; to.nF = pl->nF; to.want_block = 0; to.want_kurt = 0; to.dbg = dbg;
  for (int solo = 0; solo < 2; solo++) {
    if (solo) pl->td_smem = 120 * 1024;   // one CTA per SM: phase latencies without a co-resident CTA
    for (int rep = 0; rep < 2; rep++) { launch_td<int16_t>(pl, b, pcm, to, 0); cudaDeviceSynchronize(); }
    long long hd[16]; cudaMemcpy(hd, dbg, 128, cudaMemcpyDeviceToHost);
    printf(solo ? "--- 1 CTA/SM\n" : "--- default occupancy\n");
    const char* nm[6] = {"stage(load+cvt)", "smem->regs f64", "forward iir", "backward iir", "f64->f32 xf", "crest"};
    for (int i = 0; i < 6; i++) printf("%-16s %lld cycles\n", nm[i], hd[i + 1] - hd[i]);
    long long gs[64]; cudaMemcpyFromSymbol(gs, apt::g_stamp, sizeof(gs));
    printf("  fwd: direct %lld, scan %lld, exchange %lld, correction %lld\n", gs[0] - hd[2], gs[1] - gs[0], gs[2] - gs[1], gs[3] - gs[2]);
    printf("  bwd: direct %lld, scan %lld, exchange %lld, correction %lld\n", gs[10] - hd[3], gs[11] - gs[10], gs[12] - gs[11], gs[13] - gs[12]);
  }
    {   // STFT kernel phases (same batch)
        StftOut so; memset(&so, 0, sizeof(so)); so.P_band = pl->d_Pband.p; so.freqs = pl->d_freqs.p; so.nF = pl->nF;
        for (int rep = 0; rep < 2; rep++) { launch_stft<double, int16_t>(pl, b, pcm, so, 0); cudaDeviceSynchronize(); }
        long long gs[64]; cudaMemcpyFromSymbol(gs, apt::g_stamp, sizeof(gs));
        printf("stft256<f64>: stage %lld, pass A %lld, pass B + power %lld, plane writes %lld cycles\n", gs[21] - gs[20], gs[22] - gs[21], gs[23] - gs[22], gs[24] - gs[23]);
        for (int rep = 0; rep < 2; rep++) { launch_stft<float, int16_t>(pl, b, pcm, so, 0); cudaDeviceSynchronize(); }
        cudaMemcpyFromSymbol(gs, apt::g_stamp, sizeof(gs));
        printf("stft256<f32>: stage %lld, pass A %lld, pass B + power %lld, plane writes %lld cycles\n", gs[21] - gs[20], gs[22] - gs[21], gs[23] - gs[22], gs[24] - gs[23]);
    }
    printf("chunk=%d lb_max=%d smem=%zu tiles=%lld\n", pl->tdt.chunk, pl->tdt.lb_max, pl->td_smem, (long long)pl->td_tile_off[n_clips]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
