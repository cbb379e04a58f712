// aerobulk.cpp -- C++ API `aerobulk::model` on top of the C ABI of libaerobulk_gpu.so.
// Behaviour follows the reference's src/aerobulk.cpp:22-138 (enum -> string, equal-size
// assertion, outputs resized to the input length, scalars passed by address), with the
// two latent ABI bugs of the original closed: sizes are narrowed to int before the
// varargs call, and the skin flag crosses the boundary as one byte.
#pragma GCC visibility push(default)
#include "../../include/aerobulk.hpp"
#include "../../include/aerobulk_gpu.h"
#pragma GCC visibility pop

namespace aerobulk
{

std::string algorithm_to_string(algorithm algo)
{
    switch (algo) {
    case algorithm::OTHER:    return "other";
    case algorithm::COARE3p0: return "coare3p0";
    case algorithm::COARE3p6: return "coare3p6";
    case algorithm::NCAR:     return "ncar";
    case algorithm::ECMWF:    return "ecmwf";
    case algorithm::ANDREAS:  return "andreas";
    }
    return "unknown";
}

int check_sizes(int count, ...)
{
    va_list ap;
    va_start(ap, count);
    const int first = va_arg(ap, int);
    for (int i = 1; i < count; ++i) {
        const int other = va_arg(ap, int);
        assert(first == other);
        (void)other;
    }
    va_end(ap);
    return first;
}

static inline int isz(const std::vector<double> &v) { return static_cast<int>(v.size()); }

void model(const int jt, const int Nt, algorithm algo, double zt, double zu,
           const std::vector<double> &sst, const std::vector<double> &t_zt, const std::vector<double> &hum_zt,
           const std::vector<double> &U_zu, const std::vector<double> &V_zu, const std::vector<double> &slp,
           std::vector<double> &QL, std::vector<double> &QH, std::vector<double> &Tau_x, std::vector<double> &Tau_y,
           std::vector<double> &Evap, const int Niter, const bool l_use_skin,
           const std::vector<double> &rad_sw, const std::vector<double> &rad_lw, std::vector<double> &T_s)
{
    const std::string calgo = algorithm_to_string(algo);
    const int l = static_cast<int>(calgo.size());
    const int m = check_sizes(8, isz(sst), isz(t_zt), isz(hum_zt), isz(U_zu), isz(V_zu), isz(slp), isz(rad_sw), isz(rad_lw));
    for (std::vector<double> *out : {&QL, &QH, &Tau_x, &Tau_y, &Evap, &T_s}) out->resize(m);
    aerobulk_cxx_skin(&jt, &Nt, calgo.c_str(), &zt, &zu, sst.data(), t_zt.data(), hum_zt.data(), U_zu.data(),
                      V_zu.data(), slp.data(), QL.data(), QH.data(), Tau_x.data(), Tau_y.data(), Evap.data(),
                      &Niter, &l_use_skin, rad_sw.data(), rad_lw.data(), T_s.data(), &l, &m);
}

void model(const int jt, const int Nt, algorithm algo, double zt, double zu,
           const std::vector<double> &sst, const std::vector<double> &t_zt, const std::vector<double> &hum_zt,
           const std::vector<double> &U_zu, const std::vector<double> &V_zu, const std::vector<double> &slp,
           std::vector<double> &QL, std::vector<double> &QH, std::vector<double> &Tau_x, std::vector<double> &Tau_y,
           std::vector<double> &Evap, const int Niter)
{
    const std::string calgo = algorithm_to_string(algo);
    const int l = static_cast<int>(calgo.size());
    const int m = check_sizes(6, isz(sst), isz(t_zt), isz(hum_zt), isz(U_zu), isz(V_zu), isz(slp));
    for (std::vector<double> *out : {&QL, &QH, &Tau_x, &Tau_y, &Evap}) out->resize(m);
    aerobulk_cxx_no_skin(&jt, &Nt, calgo.c_str(), &zt, &zu, sst.data(), t_zt.data(), hum_zt.data(), U_zu.data(),
                         V_zu.data(), slp.data(), QL.data(), QH.data(), Tau_x.data(), Tau_y.data(), Evap.data(),
                         &Niter, &l, &m);
}

}  // namespace aerobulk
