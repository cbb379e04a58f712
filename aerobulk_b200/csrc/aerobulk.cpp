// aerobulk.cpp -- the C++ API `aerobulk::model` of include/aerobulk.hpp on top of the C ABI of libaerobulk_gpu.so.
//
// What a caller of the reference's src/aerobulk.cpp:22-138 relies on is kept: the enumerator -> name mapping, the
// equal-length assertion on the inputs, outputs resized to that length, the grid seen as (m, 1).  The fields go
// straight to aerobulk_gpu_model (no detour through the Fortran-style by-address bridge, whose two latent ABI slips --
// size_t read as int through varargs, a C++ bool read as a 4-byte LOGICAL -- therefore cannot occur here).
#pragma GCC visibility push(default)
#include "../../include/aerobulk.hpp"
#include "../../include/aerobulk_gpu.h"
#pragma GCC visibility pop

#include <initializer_list>

namespace aerobulk
{

namespace
{
const char *const kAlgorithmNames[] = {"other", "coare3p0", "coare3p6", "ncar", "ecmwf", "andreas"};

// common length of the input fields (asserted), as an int for the C ABI
int common_length(std::initializer_list<const field *> inputs)
{
    const std::size_t m = (*inputs.begin())->size();
    for (const field *f : inputs) {
        assert(f->size() == m && "aerobulk::model: input fields differ in length");
        (void)f;
    }
    return static_cast<int>(m);
}

void run(int jt, int Nt, algorithm algo, double zt, double zu, int m, const field &sst, const field &t_zt,
         const field &hum_zt, const field &U_zu, const field &V_zu, const field &slp, field &QL, field &QH, field &Tau_x,
         field &Tau_y, field &Evap, int Niter, const int *l_use_skin, const field *rad_sw, const field *rad_lw, field *T_s)
{
    for (field *out : {&QL, &QH, &Tau_x, &Tau_y, &Evap}) out->resize(m);
    if (T_s) T_s->resize(m);
    // errors follow the library's mode: fail-stop by default, like the reference's STOP
    aerobulk_gpu_model(jt, Nt, algorithm_to_string(algo).c_str(), zt, zu, m, 1, sst.data(), t_zt.data(), hum_zt.data(),
                       U_zu.data(), V_zu.data(), slp.data(), QL.data(), QH.data(), Tau_x.data(), Tau_y.data(), Evap.data(),
                       &Niter, l_use_skin, rad_sw ? rad_sw->data() : nullptr, rad_lw ? rad_lw->data() : nullptr,
                       T_s ? T_s->data() : nullptr);
}
}  // namespace

std::string algorithm_to_string(algorithm algo)
{
    const int k = static_cast<int>(algo);
    return (k >= 0 && k < 6) ? kAlgorithmNames[k] : "unknown";
}

int check_sizes(int count, ...)
{
    va_list sizes;
    va_start(sizes, count);
    int common = 0;
    for (int i = 0; i < count; ++i) {
        const int s = va_arg(sizes, int);
        if (i == 0) common = s;
        assert(s == common);
    }
    va_end(sizes);
    return common;
}

void model(const int jt, const int Nt, algorithm algo, double zt, double zu, const field &sst, const field &t_zt,
           const field &hum_zt, const field &U_zu, const field &V_zu, const field &slp, field &QL, field &QH, field &Tau_x,
           field &Tau_y, field &Evap, const int Niter, const bool l_use_skin, const field &rad_sw, const field &rad_lw,
           field &T_s)
{
    const int m = common_length({&sst, &t_zt, &hum_zt, &U_zu, &V_zu, &slp, &rad_sw, &rad_lw});
    const int skin = l_use_skin ? 1 : 0;
    run(jt, Nt, algo, zt, zu, m, sst, t_zt, hum_zt, U_zu, V_zu, slp, QL, QH, Tau_x, Tau_y, Evap, Niter, &skin, &rad_sw,
        &rad_lw, &T_s);
}

void model(const int jt, const int Nt, algorithm algo, double zt, double zu, const field &sst, const field &t_zt,
           const field &hum_zt, const field &U_zu, const field &V_zu, const field &slp, field &QL, field &QH, field &Tau_x,
           field &Tau_y, field &Evap, const int Niter)
{
    const int m = common_length({&sst, &t_zt, &hum_zt, &U_zu, &V_zu, &slp});
    run(jt, Nt, algo, zt, zu, m, sst, t_zt, hum_zt, U_zu, V_zu, slp, QL, QH, Tau_x, Tau_y, Evap, Niter, nullptr, nullptr,
        nullptr, nullptr);
}

}  // namespace aerobulk
