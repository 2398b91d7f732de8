// Host check of the 1-d-pass forms of the constant tables (nsdg_momentum_uniform.cuh: evalGaussSep, projectSep, divergenceSep,
// q2Values / q2Derivs) against the tables they replace (PSI<DG,3>, UnitOps::B, ::D1, ::D2, ::L, ::Lp): every unit vector in,
// maximum deviation out.  Built and run by tests/test_separable_tables.py (no GPU needed: the functions are __host__ __device__).
#include <cmath>
#include <cstdio>

#include "nsdg_momentum_uniform.cuh"

using namespace nsdg;

int main()
{
    double eval = 0, proj = 0, div = 0, q2 = 0;
    for (int j = 0; j < 8; ++j) {
        double c[8] = { 0 }, o[9];
        c[j] = 1;
        evalGaussSep<8>(c, o);
        for (int q = 0; q < 9; ++q)
            eval = fmax(eval, fabs(o[q] - PSI(3, j, q)));
        if (j < 6) {
            double c6[6] = { 0 };
            c6[j] = 1;
            evalGaussSep<6>(c6, o);
            for (int q = 0; q < 9; ++q)
                eval = fmax(eval, fabs(o[q] - PSI(3, j, q)));
        }
        double T[9];
        divergenceSep<0, false>(c, 1.7, 1.7 / 3, T);
        for (int k = 0; k < 9; ++k)
            div = fmax(div, fabs(T[k] - 1.7 * kUnitOps.D1[k][j]));
        divergenceSep<1, true>(c, 0.6, 0.2, T);
        for (int k = 0; k < 9; ++k)
            div = fmax(div, fabs(T[k] - 1.7 * kUnitOps.D1[k][j] - 0.6 * kUnitOps.D2[k][j]));
    }
    for (int q = 0; q < 9; ++q) {
        double r[9] = { 0 }, c8[8], c6[6];
        r[q] = 1;
        projectSep<8>(r, c8);
        projectSep<6>(r, c6);
        for (int j = 0; j < 8; ++j)
            proj = fmax(proj, fabs(c8[j] - kUnitOps.B[j][q]) / fmax(1.0, fabs(kUnitOps.B[j][q])));
        for (int j = 0; j < 6; ++j)
            proj = fmax(proj, fabs(c6[j] - kUnitOps.B[j][q]) / fmax(1.0, fabs(kUnitOps.B[j][q])));
    }
    for (int j = 0; j < 3; ++j) {
        double c[3] = { 0 }, v[3], d[3];
        c[j] = 1;
        q2Values(c[0], c[1], c[2], v[0], v[1], v[2]);
        q2Derivs(c[0], c[1], c[2], d[0], d[1], d[2]);
        for (int q = 0; q < 3; ++q) {
            q2 = fmax(q2, fabs(v[q] - kUnitOps.L[j][q]));
            q2 = fmax(q2, fabs(d[q] - kUnitOps.Lp[j][q]));
        }
    }
    printf("eval %.3e proj %.3e div %.3e q2 %.3e\n", eval, proj, div, q2);
    return (eval < 1e-15 && proj < 1e-15 && div < 1e-15 && q2 < 1e-14) ? 0 : 1;
}
