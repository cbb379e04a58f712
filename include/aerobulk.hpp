// aerobulk.hpp -- C++ interface of the B200 build of the aerobulk_model hot path.
//
// Source-compatible with the reference's include/aerobulk.hpp:13-41: same namespace,
// same `algorithm` enumeration values, same two `model` overloads (argument order,
// meaning and output resizing as in src/aerobulk.cpp:83-138).  A program written
// against the reference header recompiles unchanged and links with
// `-laerobulk_gpu` instead of `-laerobulk_cxx -laerobulk -lgfortran`.
#ifndef AEROBULK_B200_AEROBULK_HPP
#define AEROBULK_B200_AEROBULK_HPP 1

#include <cassert>
#include <cstdarg>
#include <string>
#include <vector>

namespace aerobulk
{
    // Values match the reference enumeration (include/aerobulk.hpp:13-21).
    enum class algorithm
    {
        OTHER    = 0,
        COARE3p0 = 1,
        COARE3p6 = 2,
        NCAR     = 3,
        ECMWF    = 4,
        ANDREAS  = 5
    };

    // "coare3p0", "coare3p6", "ncar", "ecmwf", "andreas"; "other" for OTHER (src/aerobulk.cpp:22-49).
    std::string algorithm_to_string(algorithm algo);

    // Asserts that `count` int-sized sizes are all equal and returns that size (src/aerobulk.cpp:52-65).
    int check_sizes(int count, ...);

    // aerobulk_model WITH radiation inputs and skin temperature output
    // (cool-skin / warm-layer schemes used when l_use_skin is true at jt == 1).
    void model(const int jt, const int Nt, algorithm algo, double zt, double zu,
               const std::vector<double> &sst, const std::vector<double> &t_zt,
               const std::vector<double> &hum_zt, const std::vector<double> &U_zu,
               const std::vector<double> &V_zu, const std::vector<double> &slp,
               std::vector<double> &QL, std::vector<double> &QH,
               std::vector<double> &Tau_x, std::vector<double> &Tau_y, std::vector<double> &Evap,
               const int Niter, const bool l_use_skin,
               const std::vector<double> &rad_sw, const std::vector<double> &rad_lw,
               std::vector<double> &T_s);

    // aerobulk_model without radiation inputs (bulk SST).
    void model(const int jt, const int Nt, algorithm algo, double zt, double zu,
               const std::vector<double> &sst, const std::vector<double> &t_zt,
               const std::vector<double> &hum_zt, const std::vector<double> &U_zu,
               const std::vector<double> &V_zu, const std::vector<double> &slp,
               std::vector<double> &QL, std::vector<double> &QH,
               std::vector<double> &Tau_x, std::vector<double> &Tau_y, std::vector<double> &Evap,
               const int Niter);
}

#endif
