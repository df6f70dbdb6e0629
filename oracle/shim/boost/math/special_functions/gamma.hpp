/* TEST INFRASTRUCTURE ONLY (oracle shim) - stands in for <boost/math/special_functions/gamma.hpp>:
 * gamma_p_inv(a, p) = x such that P(a, x) = p (regularised lower incomplete gamma), the one function
 * pcps_acquisition.cc:55 uses.  Series / continued fraction for P, then bisection + Newton in double. */
#pragma once
#include <cmath>
#include <limits>
namespace boost
{
namespace math
{
inline double shim_gamma_p(double a, double x)
{
    if (x <= 0.0) return 0.0;
    const double gln = std::lgamma(a);
    if (x < a + 1.0)
        {
            double ap = a, sum = 1.0 / a, del = sum;
            for (int n = 0; n < 10000; n++)
                {
                    ap += 1.0;
                    del *= x / ap;
                    sum += del;
                    if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
                }
            return sum * std::exp(-x + a * std::log(x) - gln);
        }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 10000; i++)
        {
            const double an = -i * (i - a);
            b += 2.0;
            d = an * d + b;
            if (std::fabs(d) < tiny) d = tiny;
            c = b + an / c;
            if (std::fabs(c) < tiny) c = tiny;
            d = 1.0 / d;
            const double del = d * c;
            h *= del;
            if (std::fabs(del - 1.0) < 1e-17) break;
        }
    return 1.0 - std::exp(-x + a * std::log(x) - gln) * h;
}
template <typename T1, typename T2>
inline double gamma_p_inv(T1 a_, T2 p_)
{
    const double a = static_cast<double>(a_), p = static_cast<double>(p_);
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return std::numeric_limits<double>::infinity();
    double lo = 0.0, hi = a + 10.0;
    while (shim_gamma_p(a, hi) < p) hi *= 2.0;
    for (int i = 0; i < 200; i++)
        {
            const double mid = 0.5 * (lo + hi);
            if (shim_gamma_p(a, mid) < p)
                lo = mid;
            else
                hi = mid;
            if (hi - lo <= 1e-15 * hi) break;
        }
    return 0.5 * (lo + hi);
}
}  // namespace math
}  // namespace boost
