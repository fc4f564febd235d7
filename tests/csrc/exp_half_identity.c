/* Host restatement of the two device sequences ab_exp_neg(-0.5 * r2) and ab_exp_neg_half(r2)
 * (alabi_b200/csrc/common.cuh): the same IEEE-754 operations in the same order (fma / mul / add are
 * correctly rounded on both machines), so equal bits here mean equal bits on the device.  The table
 * value only enters through one final fma that both sequences share.  Prints the number of
 * mismatches over random and special arguments; exit status 0 iff there is none. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static double tab[64];
static int32_t lo32(double t) { uint64_t u; memcpy(&u, &t, 8); return (int32_t)(uint32_t)u; }
static int32_t hi32(double t) { uint64_t u; memcpy(&u, &t, 8); return (int32_t)(uint32_t)(u >> 32); }
static double scale(double p, int n) {
    uint64_t u; memcpy(&u, &p, 8);
    uint32_t hi = (uint32_t)(u >> 32) + (uint32_t)((n >> 6) << 20);
    u = ((uint64_t)hi << 32) | (uint32_t)u;
    memcpy(&p, &u, 8);
    return p;
}
static double exp_neg(double x) {
    const double SHIFT = 6755399441055744.0;
    const double t = fma(x, 92.33248261689366, SHIFT);
    const int n = lo32(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -0x1.62e42fee00000p-7, x);
    r = fma(nf, -0x1.a39ef35793c76p-39, r);
    double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q *= r;
    const double tj = tab[n & 63];
    const double p = fma(tj, q, tj);
    return ((hi32(x) & 0x7fffffff) < 0x40861800) ? scale(p, n) : 0.0;
}
static double exp_neg_half(double r2) {
    const double SHIFT = 6755399441055744.0;
    const double t = fma(r2, -46.16624130844683, SHIFT);
    const int n = lo32(t);
    const double nf = t - SHIFT;
    double u = fma(nf, 0x1.62e42fee00000p-6, r2);
    u = fma(nf, 0x1.a39ef35793c76p-38, u);
    double q = fma(u, -8.3333333333333332e-03 * 0.03125, 4.1666666666666664e-02 * 0.0625);
    q = fma(q, u, -1.6666666666666666e-01 * 0.125);
    q = fma(q, u, 0.125);
    q = fma(q, u, -0.5);
    q *= u;
    const double tj = tab[n & 63];
    const double p = fma(tj, q, tj);
    return ((hi32(r2) & 0x7fffffff) < 0x40961800) ? scale(p, n) : 0.0;
}
int main(void) {
    for (int j = 0; j < 64; j++) tab[j] = exp2(j / 64.0);
    uint64_t s = 88172645463325252ull;
    long bad = 0, n = 0;
    double worst = 0.0;
    for (long i = 0; i < 20000000; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u01 = (double)(s >> 11) / 9007199254740992.0, r2;
        switch (i & 3) {
            case 0: r2 = u01 * 1500.0; break;
            case 1: r2 = u01 * 40.0; break;
            case 2: r2 = u01 * u01 * 1e-3; break;
            default: r2 = ldexp(u01, -(int)(s & 63)); break;
        }
        const double a = exp_neg(-0.5 * r2), b = exp_neg_half(r2);
        n++;
        if (memcmp(&a, &b, 8) != 0) bad++;
        const double e = exp(-0.5 * r2);
        if (e > 1e-300) { double rel = fabs(b - e) / e; if (rel > worst) worst = rel; }
    }
    const double sp[] = {0.0, 1413.9999999999998, 1414.0, 1414.0000000000002, 1e300, 5e-324, 2.0 * 0.6931471805599453 / 64.0};
    for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++) {
        const double a = exp_neg(-0.5 * sp[i]), b = exp_neg_half(sp[i]);
        n++;
        if (memcmp(&a, &b, 8) != 0) bad++;
    }
    printf("checked %ld mismatches %ld worst_rel_err_vs_libm %.3e\n", n, bad, worst);
    return bad != 0;
}
