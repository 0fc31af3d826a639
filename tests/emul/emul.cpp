// Host-side emulation of kernel math (tests only): compiles csrc/apt_math.cuh with g++ so the
// lane-level functions the CUDA kernels call can be checked on a CPU-only box.
#include "../../audio_processing_tools_b200/csrc/apt_math.cuh"
#include <vector>
using namespace apt;

template <typename T>
static void rfft256(const float* x, const double* win, double* out) {
    std::vector<T> w(256);
    std::vector<cx<T>> tw128(128), twA(128), tw256(129), ex(kExSize);
    for (int i = 0; i < 256; i++) w[i] = (T)win[i];
    for (int m = 0; m < 128; m++) tw128[m] = {(T)cos(2.0 * M_PI * m / 128.0), (T)-sin(2.0 * M_PI * m / 128.0)};
    for (int i = 0; i < 128; i++) twA[i] = tw128[((i & 7) * (i >> 3)) & 127];   // [k1][lane] layout of the kernel
    for (int k = 0; k <= 128; k++) tw256[k] = {(T)cos(2.0 * M_PI * k / 256.0), (T)-sin(2.0 * M_PI * k / 256.0)};
    auto ldx = [&](int n) { return x[n]; };
    for (int j = 0; j < 8; j++) rfft256_passA<T>(j, ldx, w.data(), twA.data(), ex.data());
    for (int t = 0; t < 8; t++)
        rfft256_passB<T>(t, ex.data(), tw256.data(), [&](int k, T re, T im) { out[2 * k] = (double)re; out[2 * k + 1] = (double)im; });
}
extern "C" {
void emul_rfft256_f64(const float* x, const double* win, double* out) { rfft256<double>(x, win, out); }
void emul_rfft256_f32(const float* x, const double* win, double* out) { rfft256<float>(x, win, out); }
void emul_log10f(const float* x, float* y, long n) {
    const float* tab = (const float*)kSvmlLog10TabHost.t;
    for (long i = 0; i < n; i++) y[i] = svml_log10f(x[i], tab);
}
void emul_log1pf(const float* x, float* y, long n) { for (long i = 0; i < n; i++) y[i] = svml_log1pf(x[i]); }
void emul_cabsf(const float* z, float* y, long n) { for (long i = 0; i < n; i++) y[i] = np_cabsf(z[2 * i], z[2 * i + 1]); }
void emul_pcm(float* y) { for (int i = -32768; i < 32768; i++) y[i + 32768] = pcm_to_f32((int16_t)i); }
float emul_pairwise_f32(const float* a, int n) { return 0.0f + np_pairwise<float>([&](int i) { return a[i]; }, 0, n); }
}

// Scans every float32 w with bit pattern in [lo, hi]: records the intervals [plateau start, window end] around every
// place where d(w) = 10 * log10f_svml(w) is smaller than d of a smaller w.  Returns the number of intervals found
// (at most cap are written as bit-pattern pairs).  Threads split the range on plateau-safe boundaries.
#include <thread>
#include <vector>
extern "C" int emul_db_monotone_scan(unsigned lo, unsigned hi, unsigned* out, int cap, int n_thr) {
    static float tab[64];
    for (int i = 0; i < 64; i++) tab[i] = apt::u2f(((const uint32_t*)apt::kSvmlLog10TabHost.t)[i]);
    struct Win { unsigned p, e; };
    std::vector<std::vector<Win>> found(n_thr);
    std::vector<std::thread> th;
    for (int t = 0; t < n_thr; t++) th.emplace_back([&, t] {
        // each thread starts 4096 floats early so that its running maximum is warmed up before its own range
        const unsigned long long span = (unsigned long long)hi - lo + 1;
        const unsigned a = (unsigned)(lo + span * t / n_thr), b = (unsigned)(lo + span * (t + 1) / n_thr - 1);
        const unsigned warm = a - lo > 4096 ? a - 4096 : lo;
        float runmax = 10.0f * apt::svml_log10f(apt::u2f(warm), tab);
        unsigned plateau = warm, start = 0;
        bool in = false;
        for (unsigned long long uu = (unsigned long long)warm + 1; uu <= b; uu++) {
            const unsigned u = (unsigned)uu;
            const float d = 10.0f * apt::svml_log10f(apt::u2f(u), tab);
            if (d < runmax) { if (!in) { in = true; start = u; } }
            else {
                if (in) { if (start >= a) found[t].push_back({plateau, u - 1}); in = false; }
                if (d > runmax) { runmax = d; plateau = u; }
            }
        }
        if (in && start >= a) found[t].push_back({plateau, b});
    });
    for (auto& x : th) x.join();
    int n = 0;
    for (auto& v : found) for (auto& w : v) { if (n < cap) { out[2 * n] = w.p; out[2 * n + 1] = w.e; } n++; }
    return n;
}
