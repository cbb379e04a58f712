// aerobulk.hpp -- C++ interface of the B200 build of the aerobulk_model hot path.
//
// Drop-in for the header of the same name in the reference (include/aerobulk.hpp:13-41): a program written against
// that header recompiles unchanged -- same namespace, same enumerators with the same values, same two `model`
// overloads (argument order, meaning and output resizing as in the reference's src/aerobulk.cpp:83-138) -- and links
// with `-laerobulk_gpu` instead of `-laerobulk_cxx -laerobulk -lgfortran`.
#ifndef AEROBULK_B200_AEROBULK_HPP
#define AEROBULK_B200_AEROBULK_HPP 1

#include <cassert>
#include <cstdarg>
#include <string>
#include <vector>

namespace aerobulk
{
    // One FP64 value per grid point (all fields of a call have the same length).
    using field = std::vector<double>;

    // Bulk algorithm selector; the numbering is part of the interface (OTHER is reserved for user-supplied schemes
    // and is rejected by model()).
    enum class algorithm : int { OTHER = 0, COARE3p0, COARE3p6, NCAR, ECMWF, ANDREAS };
    static_assert(static_cast<int>(algorithm::ANDREAS) == 5, "enumerator values are 0..5 in declaration order");

    // Name understood by the library: "coare3p0", "coare3p6", "ncar", "ecmwf", "andreas" ("other" for OTHER).
    std::string algorithm_to_string(algorithm algo);

    // Variadic helper kept for source compatibility: asserts that the `count` int sizes that follow are all equal
    // and returns that size.
    int check_sizes(int count, ...);

    // aerobulk_model WITH radiation: the cool-skin / warm-layer schemes are used when l_use_skin is true at jt == 1
    // (COARE 3.x and ECMWF only); T_s returns the skin temperature (or a copy of sst).
    void model(const int jt,            // time step, 1-based
               const int Nt,            // number of time steps of the session
               algorithm algo,
               double zt,               // height of t_zt and hum_zt [m]
               double zu,               // height of the wind [m]
               const field &sst,        // bulk sea-surface temperature [K]
               const field &t_zt,       // absolute air temperature at zt [K]
               const field &hum_zt,     // specific humidity [kg/kg], dew-point [K] or relative humidity [%] at zt
               const field &U_zu,       // zonal wind at zu [m/s]
               const field &V_zu,       // meridional wind at zu [m/s]
               const field &slp,        // sea-level pressure [Pa]
               field &QL,               // latent heat flux [W/m^2]            (outputs are resized)
               field &QH,               // sensible heat flux [W/m^2]
               field &Tau_x,            // zonal wind stress [N/m^2]
               field &Tau_y,            // meridional wind stress [N/m^2]
               field &Evap,             // evaporation [kg/m^2/s]
               const int Niter,         // iterations of the bulk algorithm
               const bool l_use_skin,
               const field &rad_sw,     // downwelling short-wave radiation [W/m^2]
               const field &rad_lw,     // downwelling long-wave radiation [W/m^2]
               field &T_s);             // skin temperature [K]

    // aerobulk_model without radiation (bulk SST, no skin schemes); arguments as above.
    void model(const int jt, const int Nt, algorithm algo, double zt, double zu,
               const field &sst, const field &t_zt, const field &hum_zt, const field &U_zu, const field &V_zu,
               const field &slp, field &QL, field &QH, field &Tau_x, field &Tau_y, field &Evap, const int Niter);
}

#endif
