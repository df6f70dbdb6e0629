/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <gnuradio/fft/fft.h> (gr-fft over FFTW3f, third-party,
 * not under /root/reference and not installed).  gr::fft::fft_complex_fwd / _rev keep the documented interface
 * (get_inbuf, get_outbuf, execute; unnormalised transforms, exponent sign -1 forward / +1 reverse).
 *
 * The transform itself is our own float32 FFT: Stockham autosort, mixed radix 4/2/3/5/7 with a generic
 * small-prime butterfly (11, 13) and Bluestein's chirp-z for lengths with larger prime factors; twiddles are
 * computed in double and rounded once.  It is the CPU transform behind the acquisition reference arm
 * (bench.py --impl reference) and behind the compiled-in-place pcps_acquisition oracle. */
#pragma once
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace shimfft
{
typedef std::complex<float> cf;

struct Stage
{
    int r;                 // radix
    int m;                 // n / r at this stage (n = remaining length)
    int s;                 // stride
    std::vector<cf> tw;    // m x (r-1) twiddles  w_p^k, k = 1..r-1
    std::vector<cf> dft;   // r x r DFT matrix for generic radices
};

class Plan
{
public:
    Plan(int n, bool forward) : d_n(n), d_fwd(forward)
    {
        int rem = n;
        std::vector<int> radices;
        auto take = [&](int r) {
            while (rem % r == 0)
                {
                    radices.push_back(r);
                    rem /= r;
                }
        };
        take(4);
        take(2);
        take(3);
        take(5);
        take(7);
        take(11);
        take(13);
        if (rem != 1)
            {
                d_bluestein = true;
                init_bluestein();
                return;
            }
        int len = n, s = 1;
        const double sgn = forward ? -1.0 : 1.0;
        for (int r : radices)
            {
                Stage st;
                st.r = r;
                st.m = len / r;
                st.s = s;
                st.tw.resize(static_cast<size_t>(st.m) * (r - 1));
                for (int p = 0; p < st.m; p++)
                    for (int k = 1; k < r; k++)
                        {
                            const double a = sgn * 2.0 * M_PI * static_cast<double>(p) * k / static_cast<double>(len);
                            st.tw[static_cast<size_t>(p) * (r - 1) + (k - 1)] = cf(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
                        }
                if (r >= 7)
                    {
                        st.dft.resize(static_cast<size_t>(r) * r);
                        for (int j = 0; j < r; j++)
                            for (int k = 0; k < r; k++)
                                {
                                    const double a = sgn * 2.0 * M_PI * static_cast<double>((j * k) % r) / r;
                                    st.dft[static_cast<size_t>(k) * r + j] = cf(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
                                }
                    }
                d_stages.push_back(std::move(st));
                len /= r;
                s *= r;
            }
        d_work.resize(n);
    }

    // out may alias in
    void execute(const cf* in, cf* out)
    {
        if (d_bluestein)
            {
                run_bluestein(in, out);
                return;
            }
        if (d_stages.empty())
            {
                if (out != in) std::memcpy(out, in, sizeof(cf) * d_n);
                return;
            }
        // ping-pong between out and work so that the last stage lands in out
        const size_t ns = d_stages.size();
        cf* a = (ns % 2 == 0) ? out : d_work.data();
        if (a != in) std::memcpy(a, in, sizeof(cf) * d_n);
        cf* b = (a == out) ? d_work.data() : out;
        for (size_t i = 0; i < ns; i++)
            {
                run_stage(d_stages[i], a, b);
                std::swap(a, b);
            }
    }
    int size() const { return d_n; }

private:
    static inline cf mul(cf a, cf b) { return cf(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real()); }
    inline cf rot90(cf a) const { return d_fwd ? cf(a.imag(), -a.real()) : cf(-a.imag(), a.real()); }  // multiply by -j (fwd) / +j (rev)

    void run_stage(const Stage& st, const cf* __restrict x, cf* __restrict y) const
    {
        const int r = st.r, m = st.m, s = st.s;
        switch (r)
            {
            case 2:
                for (int p = 0; p < m; p++)
                    {
                        const cf w = st.tw[p];
                        const cf* x0 = x + static_cast<size_t>(s) * p;
                        const cf* x1 = x + static_cast<size_t>(s) * (p + m);
                        cf* y0 = y + static_cast<size_t>(s) * (2 * p);
                        cf* y1 = y0 + s;
                        for (int q = 0; q < s; q++)
                            {
                                const cf a = x0[q], b = x1[q];
                                y0[q] = a + b;
                                y1[q] = mul(a - b, w);
                            }
                    }
                break;
            case 4:
                for (int p = 0; p < m; p++)
                    {
                        const cf w1 = st.tw[3 * static_cast<size_t>(p)], w2 = st.tw[3 * static_cast<size_t>(p) + 1], w3 = st.tw[3 * static_cast<size_t>(p) + 2];
                        const cf* x0 = x + static_cast<size_t>(s) * p;
                        const cf* x1 = x0 + static_cast<size_t>(s) * m;
                        const cf* x2 = x1 + static_cast<size_t>(s) * m;
                        const cf* x3 = x2 + static_cast<size_t>(s) * m;
                        cf* y0 = y + static_cast<size_t>(s) * (4 * p);
                        cf* y1 = y0 + s;
                        cf* y2 = y1 + s;
                        cf* y3 = y2 + s;
                        for (int q = 0; q < s; q++)
                            {
                                const cf a = x0[q], b = x1[q], c = x2[q], d = x3[q];
                                const cf apc = a + c, amc = a - c, bpd = b + d, jbmd = rot90(b - d);
                                y0[q] = apc + bpd;
                                y1[q] = mul(amc + jbmd, w1);
                                y2[q] = mul(apc - bpd, w2);
                                y3[q] = mul(amc - jbmd, w3);
                            }
                    }
                break;
            case 3:
                {
                    const float c3 = -0.5f, s3 = (d_fwd ? -1.0f : 1.0f) * 0.86602540378443864676f;
                    for (int p = 0; p < m; p++)
                        {
                            const cf w1 = st.tw[2 * static_cast<size_t>(p)], w2 = st.tw[2 * static_cast<size_t>(p) + 1];
                            const cf* x0 = x + static_cast<size_t>(s) * p;
                            const cf* x1 = x0 + static_cast<size_t>(s) * m;
                            const cf* x2 = x1 + static_cast<size_t>(s) * m;
                            cf* y0 = y + static_cast<size_t>(s) * (3 * p);
                            cf* y1 = y0 + s;
                            cf* y2 = y1 + s;
                            for (int q = 0; q < s; q++)
                                {
                                    const cf a = x0[q], b = x1[q], c = x2[q];
                                    const cf t1 = b + c;
                                    const cf t2 = a + c3 * t1;
                                    const cf d = b - c;
                                    const cf t3 = cf(-s3 * d.imag(), s3 * d.real());  // j*s3*(b-c)
                                    y0[q] = a + t1;
                                    y1[q] = mul(t2 + t3, w1);
                                    y2[q] = mul(t2 - t3, w2);
                                }
                        }
                    break;
                }
            case 5:
                {
                    const float sg = d_fwd ? -1.0f : 1.0f;
                    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
                    const float s1 = sg * 0.95105651629515357212f, s2 = sg * 0.58778525229247312917f;
                    for (int p = 0; p < m; p++)
                        {
                            const cf* tw = &st.tw[4 * static_cast<size_t>(p)];
                            const cf* x0 = x + static_cast<size_t>(s) * p;
                            const size_t sm = static_cast<size_t>(s) * m;
                            cf* y0 = y + static_cast<size_t>(s) * (5 * p);
                            for (int q = 0; q < s; q++)
                                {
                                    const cf a0 = x0[q], a1 = x0[q + sm], a2 = x0[q + 2 * sm], a3 = x0[q + 3 * sm], a4 = x0[q + 4 * sm];
                                    const cf t1 = a1 + a4, t2 = a2 + a3, t3 = a1 - a4, t4 = a2 - a3;
                                    const cf m1 = a0 + c1 * t1 + c2 * t2;
                                    const cf m2 = a0 + c2 * t1 + c1 * t2;
                                    const cf u1 = s1 * t3 + s2 * t4;
                                    const cf u2 = s2 * t3 - s1 * t4;
                                    const cf ju1 = cf(-u1.imag(), u1.real()), ju2 = cf(-u2.imag(), u2.real());
                                    y0[q] = a0 + t1 + t2;
                                    y0[q + s] = mul(m1 + ju1, tw[0]);
                                    y0[q + 2 * static_cast<size_t>(s)] = mul(m2 + ju2, tw[1]);
                                    y0[q + 3 * static_cast<size_t>(s)] = mul(m2 - ju2, tw[2]);
                                    y0[q + 4 * static_cast<size_t>(s)] = mul(m1 - ju1, tw[3]);
                                }
                        }
                    break;
                }
            default:
                {
                    // generic small radix (7, 11, 13): r x r matrix product
                    std::vector<cf> mat;
                    const cf* D;
                    if (st.dft.empty())
                        {
                            mat.resize(static_cast<size_t>(r) * r);
                            const double sgn = d_fwd ? -1.0 : 1.0;
                            for (int j = 0; j < r; j++)
                                for (int k = 0; k < r; k++)
                                    {
                                        const double a = sgn * 2.0 * M_PI * static_cast<double>((j * k) % r) / r;
                                        mat[static_cast<size_t>(k) * r + j] = cf(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
                                    }
                            D = mat.data();
                        }
                    else
                        D = st.dft.data();
                    const size_t sm = static_cast<size_t>(s) * m;
                    cf a[16];
                    for (int p = 0; p < m; p++)
                        {
                            const cf* tw = &st.tw[static_cast<size_t>(r - 1) * p];
                            const cf* x0 = x + static_cast<size_t>(s) * p;
                            cf* y0 = y + static_cast<size_t>(s) * (static_cast<size_t>(r) * p);
                            for (int q = 0; q < s; q++)
                                {
                                    for (int j = 0; j < r; j++) a[j] = x0[q + j * sm];
                                    for (int k = 0; k < r; k++)
                                        {
                                            cf acc = a[0];
                                            const cf* row = D + static_cast<size_t>(k) * r;
                                            for (int j = 1; j < r; j++) acc += mul(a[j], row[j]);
                                            y0[q + static_cast<size_t>(k) * s] = (k == 0) ? acc : mul(acc, tw[k - 1]);
                                        }
                                }
                        }
                }
            }
    }

    // ---- Bluestein: X[k] = conj-chirp[k] * sum_n (x[n] chirp[n]) * conj-chirp... via circular convolution of length M
    void init_bluestein()
    {
        int M = 1;
        auto smooth = [](int v) {
            for (int f : {2, 3, 5}) while (v % f == 0) v /= f;
            return v == 1;
        };
        M = 2 * d_n - 1;
        while (!smooth(M)) M++;
        d_M = M;
        d_sub_f.reset(new Plan(M, true));
        d_sub_r.reset(new Plan(M, false));
        d_chirp.resize(d_n);
        const double sgn = d_fwd ? -1.0 : 1.0;
        for (int i = 0; i < d_n; i++)
            {
                const long long i2 = (static_cast<long long>(i) * i) % (2LL * d_n);
                const double a = sgn * M_PI * static_cast<double>(i2) / d_n;
                d_chirp[i] = cf(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
            }
        std::vector<cf> b(M, cf(0.f, 0.f));
        b[0] = std::conj(d_chirp[0]);
        for (int i = 1; i < d_n; i++) b[i] = b[M - i] = std::conj(d_chirp[i]);
        d_bspec.resize(M);
        d_sub_f->execute(b.data(), d_bspec.data());
        d_ba.resize(M);
        d_bb.resize(M);
    }
    void run_bluestein(const cf* in, cf* out)
    {
        for (int i = 0; i < d_n; i++) d_ba[i] = mul(in[i], d_chirp[i]);
        for (int i = d_n; i < d_M; i++) d_ba[i] = cf(0.f, 0.f);
        d_sub_f->execute(d_ba.data(), d_bb.data());
        for (int i = 0; i < d_M; i++) d_bb[i] = mul(d_bb[i], d_bspec[i]);
        d_sub_r->execute(d_bb.data(), d_ba.data());
        const float inv = 1.0f / static_cast<float>(d_M);
        for (int i = 0; i < d_n; i++) out[i] = mul(d_ba[i], d_chirp[i]) * inv;
    }

    int d_n;
    bool d_fwd;
    std::vector<Stage> d_stages;
    std::vector<cf> d_work;
    bool d_bluestein{false};
    int d_M{0};
    std::unique_ptr<Plan> d_sub_f, d_sub_r;
    std::vector<cf> d_chirp, d_bspec, d_ba, d_bb;
};
}  // namespace shimfft

namespace gr
{
namespace fft
{
template <bool FORWARD>
class fft_complex_shim
{
public:
    explicit fft_complex_shim(int fft_size, int /*nthreads*/ = 1) : d_plan(fft_size, FORWARD), d_in(fft_size), d_out(fft_size) {}
    std::complex<float>* get_inbuf() { return d_in.data(); }
    std::complex<float>* get_outbuf() { return d_out.data(); }
    int inbuf_length() const { return d_plan.size(); }
    int outbuf_length() const { return d_plan.size(); }
    void execute() { d_plan.execute(d_in.data(), d_out.data()); }

private:
    shimfft::Plan d_plan;
    std::vector<std::complex<float>> d_in, d_out;
};
typedef fft_complex_shim<true> fft_complex_fwd;
typedef fft_complex_shim<false> fft_complex_rev;
}  // namespace fft
}  // namespace gr
