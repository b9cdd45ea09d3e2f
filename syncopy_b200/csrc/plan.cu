// FFT plan cache (host side).  All tables are computed in double precision and rounded once.
#include "plan.cuh"
#include "common.cuh"

#include <cmath>
#include <complex>
#include <map>
#include <mutex>
#include <vector>

namespace spyb {

namespace {
std::mutex g_mu;
std::map<int, FftPlan*> g_plans;
std::map<int, FftPlanD*> g_plans_d;
std::vector<void*> g_allocs;

const double kPi = 3.14159265358979323846264338327950288;

int ilog2(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

// iterative radix-2 FFT in double, host only (table generation)
void host_fft(std::vector<std::complex<double>>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * kPi * double(k) / double(len);
                std::complex<double> w(std::cos(ang), std::sin(ang));
                std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
        }
    }
}

template <typename T2>
bool upload(const std::vector<T2>& h, const T2** dev) {
    void* d = nullptr;
    if (h.empty()) { *dev = nullptr; return true; }
    if (cudaMalloc(&d, h.size() * sizeof(T2)) != cudaSuccess) return false;
    if (cudaMemcpy(d, h.data(), h.size() * sizeof(T2), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    g_allocs.push_back(d);
    *dev = static_cast<const T2*>(d);
    return true;
}

// pass twiddles in execution order: radix-16 passes with NS = 16^I, then the remainder pass
std::vector<float2> make_pass_twiddles(int log2n) {
    std::vector<float2> tw;
    const int q16 = log2n / 4, rlast = 1 << (log2n % 4);
    long long ns = 1;
    auto emit = [&](int R) {
        if (ns > 1) {
            for (int r = 1; r < R; ++r)
                for (long long k = 0; k < ns; ++k) {
                    // exponent k*r / (ns*R), reduced exactly in integers
                    const long long den = ns * R;
                    const long long num = (k * r) % den;
                    const double ang = -2.0 * kPi * double(num) / double(den);
                    tw.push_back(make_float2(float(std::cos(ang)), float(std::sin(ang))));
                }
        }
        ns *= R;
    };
    for (int i = 0; i < q16; ++i) emit(16);
    if (rlast > 1) emit(rlast);
    return tw;
}

// twiddles of the in-place decimation-in-frequency passes (mtm_dif.cu), in execution order: the pass of radix R at
// butterfly distance `stride` multiplies output q of the butterfly at offset o by W_{R*stride}^{o*q}, stored at
// [(q-1)*stride + o]; the last pass (stride 1) needs none.
std::vector<float2> make_dif_twiddles(int log2n) {
    std::vector<float2> tw;
    const int q16 = log2n / 4;
    long long stride = 1LL << log2n;
    for (int i = 0; i < q16; ++i) {
        stride /= 16;
        if (stride > 1) {
            const long long den = 16 * stride;
            for (int q = 1; q < 16; ++q)
                for (long long o = 0; o < stride; ++o) {
                    const long long num = (o * q) % den;
                    const double ang = -2.0 * kPi * double(num) / double(den);
                    tw.push_back(make_float2(float(std::cos(ang)), float(std::sin(ang))));
                }
        }
    }
    return tw;
}
}  // namespace

const FftPlan* get_fft_plan(int n_dft) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_plans.find(n_dft);
    if (it != g_plans.end()) return it->second;
    if (n_dft < 1) { fail("FFT length must be >= 1 (got %d)", n_dft); return nullptr; }

    FftPlan* pl = new FftPlan();
    pl->n_dft = n_dft;
    const bool pow2 = (n_dft & (n_dft - 1)) == 0;
    if (pow2 && n_dft >= 16) {
        pl->log2n = ilog2(n_dft);
        pl->bluestein = false;
    } else {
        int m = 16;
        while (m < 2 * n_dft - 1) m <<= 1;
        pl->log2n = ilog2(m);
        pl->bluestein = true;
    }
    if (pl->log2n > 14) {
        fail("FFT length %d not supported by the shared-memory engine (needs a block FFT of 2^%d > 2^14)",
             n_dft, pl->log2n);
        delete pl;
        return nullptr;
    }
    if (!upload(make_pass_twiddles(pl->log2n), &pl->tw)) { fail("plan upload failed"); delete pl; return nullptr; }
    if (!pl->bluestein && !upload(make_dif_twiddles(pl->log2n), &pl->tw_dif)) {
        fail("plan upload failed"); delete pl; return nullptr;
    }

    if (pl->bluestein) {
        const int n = n_dft, M = 1 << pl->log2n;
        std::vector<std::complex<double>> b(n), bw(M, 0.0);
        std::vector<float2> chirp(n), bhat(M);
        for (int i = 0; i < n; ++i) {
            const long long q = (1LL * i * i) % (2LL * n);
            const double ang = kPi * double(q) / double(n);
            b[i] = std::complex<double>(std::cos(ang), std::sin(ang));
            chirp[i] = make_float2(float(b[i].real()), float(b[i].imag()));
        }
        bw[0] = b[0];
        for (int i = 1; i < n; ++i) { bw[i] = b[i]; bw[M - i] = b[i]; }
        host_fft(bw);
        for (int i = 0; i < M; ++i)
            bhat[i] = make_float2(float(bw[i].real() / M), float(bw[i].imag() / M));
        if (!upload(chirp, &pl->chirp) || !upload(bhat, &pl->bhat)) {
            fail("plan upload failed"); delete pl; return nullptr;
        }
    }
    g_plans[n_dft] = pl;
    return pl;
}

const FftPlanD* get_fft_plan_d(int n) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_plans_d.find(n);
    if (it != g_plans_d.end()) return it->second;
    if (n < 2 || (n & (n - 1))) { fail("double FFT length must be a power of two >= 2 (got %d)", n); return nullptr; }
    FftPlanD* pl = new FftPlanD();
    pl->n = n;
    pl->log2n = ilog2(n);
    std::vector<double2> tw(n / 2);
    for (int k = 0; k < n / 2; ++k) {
        const double ang = -2.0 * kPi * double(k) / double(n);
        tw[k] = make_double2(std::cos(ang), std::sin(ang));
    }
    if (!upload(tw, &pl->tw)) { fail("plan upload failed"); delete pl; return nullptr; }
    g_plans_d[n] = pl;
    return pl;
}

void free_all_plans() {
    std::lock_guard<std::mutex> lock(g_mu);
    for (void* p : g_allocs) cudaFree(p);
    g_allocs.clear();
    for (auto& kv : g_plans) delete kv.second;
    for (auto& kv : g_plans_d) delete kv.second;
    g_plans.clear();
    g_plans_d.clear();
}

}  // namespace spyb
