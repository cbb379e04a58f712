// Caller of the C++ API `aerobulk::model` (include/aerobulk.hpp), modelled on what the reference's
// src/tests/example_call_aerobulk.cpp exercises: 2 points (unstable / stable air), the five algorithms,
// skin schemes on for COARE 3.x and ECMWF.  Prints one machine-readable line per algorithm:
//   <algo> QH0 QH1 QL0 QL1 Evap0 Evap1 Tau_x0 Tau_x1 Tau_y0 Tau_y1 T_s0 T_s1
#include <cstdio>
#include <vector>

#include "aerobulk.hpp"

int main()
{
    const double zt = 2., zu = 10.;
    const int nbiter = 8;
    const std::vector<double> sst = {273.15 + 22., 273.15 + 22.}, t_zt = {273.15 + 20., 273.15 + 25.};
    const std::vector<double> q_zt = {0.012, 0.012}, U = {4., 4.}, V = {9., 9.}, slp = {101000., 101000.};
    const std::vector<double> rsw = {0., 0.}, rlw = {350., 350.};
    std::vector<double> QL, QH, Tx, Ty, E, Ts;

    using aerobulk::algorithm;
    const algorithm algos[5] = {algorithm::COARE3p0, algorithm::COARE3p6, algorithm::ECMWF, algorithm::NCAR, algorithm::ANDREAS};
    for (algorithm a : algos) {
        const bool skin = (a == algorithm::COARE3p0 || a == algorithm::COARE3p6 || a == algorithm::ECMWF);
        if (skin) {
            aerobulk::model(1, 1, a, zt, zu, sst, t_zt, q_zt, U, V, slp, QL, QH, Tx, Ty, E, nbiter, true, rsw, rlw, Ts);
        } else {
            aerobulk::model(1, 1, a, zt, zu, sst, t_zt, q_zt, U, V, slp, QL, QH, Tx, Ty, E, nbiter);
            Ts = sst;
        }
        if (QL.size() != 2 || Ts.size() != 2) return 3;
        std::printf("RESULT %s %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n",
                    aerobulk::algorithm_to_string(a).c_str(), QH[0], QH[1], QL[0], QL[1], E[0], E[1], Tx[0], Tx[1],
                    Ty[0], Ty[1], Ts[0], Ts[1]);
    }
    return 0;
}
