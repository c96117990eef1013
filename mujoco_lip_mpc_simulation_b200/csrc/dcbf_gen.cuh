// dcbf_gen.cuh -- scenario generation next to the solver.
//
// The reference draws one obstacle field per run with Python's global `random` (rand_obs.py:31-81): circles are rejection
// sampled against two keep-out discs (start and goal) and against each other, every second one becomes an ellipse in 'mix'
// mode.  The loop there has no exit when a field cannot be completed.  Here the same rule runs for F fields at once with a
// counter-based generator (Philox4x32-10: every draw is a pure function of (seed, stream, field, draw number), so the batch
// does not depend on the launch geometry and a numpy mirror -- oracle/scenario_gen.py -- reproduces it bit for bit), a
// field that stalls is restarted, and the number of restarts is bounded: a field that cannot be built reports draws = -1.
//
// Arithmetic that decides acceptance is written with explicit round-to-nearest multiplies and adds (no FMA contraction), so
// host, device and numpy agree exactly.  The file compiles for the host (tests/hostsim).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DCBF_GHD __host__ __device__ inline
#else
#define DCBF_GHD inline
#endif

namespace dcbf {
namespace gen {

#if defined(__CUDA_ARCH__)
#define GEN_MUL(a, b) __dmul_rn((a), (b))
#define GEN_ADD(a, b) __dadd_rn((a), (b))
#define GEN_SUB(a, b) __dsub_rn((a), (b))
#else
#define GEN_MUL(a, b) ((a) * (b))
#define GEN_ADD(a, b) ((a) + (b))
#define GEN_SUB(a, b) ((a) - (b))
#endif

enum { STREAM_FIELD = 1, STREAM_MIX = 2, STREAM_POS = 3, STREAM_STATE = 4 };

DCBF_GHD void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

// Philox4x32 with 10 rounds (Salmon et al., SC'11)
DCBF_GHD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c[0], h0, l0);
        mulhilo(0xCD9E8D57u, c[2], h1, l1);
        const uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// two uniforms in [0, 1) with 53 random bits each: block `blk` of entity `idx` on `stream`
DCBF_GHD void uniform2(uint64_t seed, uint32_t stream, uint32_t idx, uint32_t blk, double &u0, double &u1) {
    uint32_t c[4] = {idx, blk, stream, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c[0] << 32) | c[1], b = ((uint64_t)c[2] << 32) | c[3];
    u0 = (double)(a >> 11) * 0x1.0p-53;
    u1 = (double)(b >> 11) * 0x1.0p-53;
}

// two decimals, like round(x, 2) of rand_obs.py up to ties
DCBF_GHD double round2(double x) { return rint(GEN_MUL(x, 100.0)) / 100.0; }

struct FieldSpec {
    int num, mix;
    double margin, radius, half_gap, safe_dis;
    int stall, max_restarts;
};

// (x - ox)^2 + (y - oy)^2 - (r + or + 2 half_gap)^2 >= 0   (rand_obs.py:39-41, same operation order)
DCBF_GHD bool clear_of(double x, double y, double r, double ox, double oy, double orad, double half_gap) {
    const double dx = GEN_SUB(x, ox), dy = GEN_SUB(y, oy);
    const double s = GEN_ADD(GEN_ADD(r, orad), GEN_MUL(2.0, half_gap));
    return GEN_SUB(GEN_ADD(GEN_MUL(dx, dx), GEN_MUL(dy, dy)), GEN_MUL(s, s)) >= 0.0;
}

#define DCBF_GEN_MAX_OBS 32

// One obstacle field.  cir[Kc][3] / elp[Ke][5] receive the inflated obstacles (radius and semi-axes + safe_dis), Kc = num
// ('cir') or ceil(num / 2) ('mix'), Ke = 0 or floor(num / 2).  Returns the number of candidates drawn, -1 if the field could
// not be completed within max_restarts restarts (outputs are then NaN).
DCBF_GHD int make_field(const FieldSpec &S, uint64_t seed, uint32_t f, double *cir, double *elp) {
    double px[DCBF_GEN_MAX_OBS], py[DCBF_GEN_MAX_OBS], pr[DCBF_GEN_MAX_OBS];
    const int Kc = S.mix ? (S.num + 1) / 2 : S.num, Ke = S.mix ? S.num / 2 : 0;
    uint32_t q = 0;   // candidates drawn so far (never reset: a restart continues the stream)
    bool done = false;
    for (int attempt = 0; attempt <= S.max_restarts && !done; attempt++) {
        int placed = 0;
        for (int tries = 0; tries < S.stall && placed < S.num; tries++, q++) {
            double u0, u1, u2, u3;
            uniform2(seed, STREAM_FIELD, f, 2u * q, u0, u1);
            uniform2(seed, STREAM_FIELD, f, 2u * q + 1u, u2, u3);
            const double x = round2(GEN_MUL(S.margin, u0)), y = round2(GEN_MUL(S.margin, u1));
            const double r = round2(GEN_ADD(GEN_MUL(GEN_SUB(S.radius, 0.35), u2), 0.35));
            bool ok = clear_of(x, y, r, 10.0, 10.0, 0.3, S.half_gap) && clear_of(x, y, r, 0.0, 0.0, 1.0, S.half_gap);
            for (int j = 0; j < placed && ok; j++) ok = clear_of(x, y, r, px[j], py[j], pr[j], S.half_gap);
            if (ok) { px[placed] = x; py[placed] = y; pr[placed] = r; placed++; }
        }
        done = placed == S.num;
    }
    if (!done) {
        for (int i = 0; i < 3 * Kc; i++) cir[i] = NAN;
        for (int i = 0; i < 5 * Ke; i++) elp[i] = NAN;
        return -1;
    }
    for (int i = 0; i < S.num; i++) {
        if (!S.mix || (i & 1) == 0) {
            double *o = cir + 3 * (S.mix ? i / 2 : i);
            o[0] = px[i]; o[1] = py[i]; o[2] = GEN_ADD(pr[i], S.safe_dis);
        } else {
            // rand_obs.py:64-67: b in [a/2, a), phi a whole number of degrees in [0, 180]
            double u0, u1;
            uniform2(seed, STREAM_MIX, f, (uint32_t)i, u0, u1);
            const double a = pr[i], ha = GEN_MUL(a, 0.5);
            const double b = round2(GEN_ADD(GEN_MUL(ha, u0), ha));
            const double deg = floor(GEN_MUL(u1, 181.0));
            const double phi = round2(GEN_MUL(deg, 3.141592653589793) / 180.0);
            double *o = elp + 5 * (i / 2);
            o[0] = px[i]; o[1] = py[i]; o[2] = GEN_ADD(a, S.safe_dis); o[3] = GEN_ADD(b, S.safe_dis); o[4] = phi;
        }
    }
    return (int)q;
}

// min over the obstacles of one prepared field of the level-set value at (x, y)
DCBF_GHD double field_clearance(const double *cir_rec, int Kc, int cir_stride, const double *elp_rec, int Ke, int elp_stride,
                                double x, double y) {
    double h = INFINITY;
    for (int j = 0; j < Kc; j++) {
        const double *o = cir_rec + cir_stride * j;
        const double dx = x - o[0], dy = y - o[1];
        const double v = dx * dx + dy * dy - o[2];
        h = v < h ? v : h;
    }
    for (int j = 0; j < Ke; j++) {
        const double *o = elp_rec + elp_stride * j;
        const double dx = x - o[0], dy = y - o[1];
        const double v = o[2] * dx * dx + o[3] * dx * dy + o[4] * dy * dy - o[5];
        h = v < h ? v : h;
    }
    return h;
}

struct StateSpec {
    int dd;            // 1: differential drive layout (x, y, theta)
    double gx, gy;     // goal
    double span;       // start positions are uniform in [0, span)^2
    double clearance;  // accepted when every level set is >= clearance at the start
    double jitter;     // heading = bearing to the goal + U(-jitter, jitter)
    double vbx_lo, vbx_hi, vby_lo, vby_hi;   // body-frame velocity ranges; the lateral sign is -leg
    int max_attempts;
};

// One start state.  x0[5|3], goal[2], warm[15|6] (cold start = [x0, x0, x0]; dd: (0.8, 0) three times), last_u[2] (dd).
// Returns the number of positions tried, -1 when none of max_attempts candidates was clear (x0 is NaN then).
template <class SinCos, class Atan2>
DCBF_GHD int make_state(const StateSpec &S, uint64_t seed, uint32_t b, const double *cir_rec, int Kc, int cir_stride,
                        const double *elp_rec, int Ke, int elp_stride, double *x0, double *goal, int32_t *leg, double *warm,
                        double *last_u, SinCos sincos_fn, Atan2 atan2_fn) {
    double px = NAN, py = NAN;
    int used = -1;
    for (int a = 0; a < S.max_attempts; a++) {
        double u0, u1;
        uniform2(seed, STREAM_POS, b, (uint32_t)a, u0, u1);
        const double x = GEN_MUL(S.span, u0), y = GEN_MUL(S.span, u1);
        if (field_clearance(cir_rec, Kc, cir_stride, elp_rec, Ke, elp_stride, x, y) >= S.clearance) {
            px = x; py = y; used = a + 1;
            break;
        }
    }
    double u0, u1, u2, u3;
    uniform2(seed, STREAM_STATE, b, 0u, u0, u1);
    uniform2(seed, STREAM_STATE, b, 1u, u2, u3);
    const int lg = u1 < 0.5 ? 1 : -1;
    const double theta = atan2_fn(S.gy - py, S.gx - px) + GEN_ADD(-S.jitter, GEN_MUL(GEN_MUL(2.0, S.jitter), u0));
    if (goal) { goal[0] = S.gx; goal[1] = S.gy; }
    if (leg) *leg = lg;
    if (S.dd) {
        if (x0) { x0[0] = px; x0[1] = py; x0[2] = theta; }
        if (warm) for (int k = 0; k < 3; k++) { warm[2 * k] = 0.8; warm[2 * k + 1] = 0.0; }
        if (last_u) { last_u[0] = 0.8; last_u[1] = 0.0; }
        return used;
    }
    const double vbx = GEN_ADD(S.vbx_lo, GEN_MUL(GEN_SUB(S.vbx_hi, S.vbx_lo), u2));
    const double vby = -(double)lg * GEN_ADD(S.vby_lo, GEN_MUL(GEN_SUB(S.vby_hi, S.vby_lo), u3));
    double sn, cs;
    sincos_fn(theta, &sn, &cs);
    const double vx = cs * vbx - sn * vby, vy = sn * vbx + cs * vby;
    const double st[5] = {px, py, vx, vy, theta};
    if (x0) for (int i = 0; i < 5; i++) x0[i] = st[i];
    if (warm) for (int k = 0; k < 3; k++) for (int i = 0; i < 5; i++) warm[5 * k + i] = st[i];
    return used;
}

}  // namespace gen
}  // namespace dcbf
