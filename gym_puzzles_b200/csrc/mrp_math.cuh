// mrp_math.cuh — float32 2-D math for the sm_100a MultiRobotPuzzle kernels.
//
// Every helper spells out the same operation order as Box2D 2.3.x's b2Math.h
//  so that results are bit-identical to an x86-64 build of
// Box2D, which never contracts mul+add: compile with -fmad=false (nvcc) /
// -ffp-contract=off (g++ host build used only by tests/emu).
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MRP_HD __host__ __device__ __forceinline__
#define MRP_HDN __host__ __device__ __noinline__
#else
#define MRP_HD inline
#define MRP_HDN
#endif

namespace mrp {

constexpr float kPi = 3.14159265359f;
constexpr float kLinearSlop = 0.005f;
constexpr float kPolygonRadius = 2.0f * kLinearSlop;
constexpr float kAabbExtension = 0.1f;
constexpr float kAabbMultiplier = 2.0f;
constexpr int kMaxSubSteps = 8;
constexpr float kMaxLinearCorrection = 0.2f;
constexpr float kMaxTranslation = 2.0f;
constexpr float kMaxTranslationSquared = kMaxTranslation * kMaxTranslation;
constexpr float kMaxRotation = 0.5f * kPi;
constexpr float kMaxRotationSquared = kMaxRotation * kMaxRotation;
constexpr float kBaumgarte = 0.2f;
constexpr float kToiBaumgarte = 0.75f;
constexpr float kEps = FLT_EPSILON;
constexpr int kToiMaxPushBack = 16;  // pybox2d b2_maxPolygonVertices (SURVEY.md A.12 fork 3)

struct V2 {
    float x, y;
};
MRP_HD V2 mk(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
MRP_HD V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
MRP_HD V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
MRP_HD V2 operator-(V2 a) { return mk(-a.x, -a.y); }
MRP_HD V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
MRP_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
MRP_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
MRP_HD V2 crossVS(V2 a, float s) { return mk(s * a.y, -s * a.x); }
MRP_HD V2 crossSV(float s, V2 a) { return mk(-s * a.y, s * a.x); }
MRP_HD float fmin2(float a, float b) { return a < b ? a : b; }  // b2Min
MRP_HD float fmax2(float a, float b) { return a > b ? a : b; }  // b2Max
MRP_HD float clampf(float a, float lo, float hi) { return fmax2(lo, fmin2(a, hi)); }
MRP_HD float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
MRP_HD V2 normalized(V2 a) {  // b2Vec2::Normalize
    float len = length(a);
    if (len < kEps) return a;
    float inv = 1.0f / len;
    return mk(a.x * inv, a.y * inv);
}

struct Rot {
    float s, c;
};
struct Xf {
    V2 p;
    Rot q;
};

// b2Rot::Set with correctly-rounded sin/cos (see DESIGN.md "sincos"): evaluated in
// float64 and rounded once, on the device with CUDA's sincos(double).  Not inlined: the float64 sincos expands to
// ~400 SASS instructions and the per-env kernels call it from many places; one shared copy keeps their hot code inside the
// instruction cache (k_post spent 24 % of its stall samples waiting on instruction fetch).
MRP_HDN Rot rot_set(float a) {
    Rot r;
#if defined(__CUDA_ARCH__)
    double sd, cd;
    ::sincos((double)a, &sd, &cd);
    r.s = (float)sd;
    r.c = (float)cd;
#else
    r.s = (float)::sin((double)a);
    r.c = (float)::cos((double)a);
#endif
    return r;
}
// inlined form for the persistent position solver, whose one call site sits in its hot loop
MRP_HD Rot rot_set_inline(float a) {
    Rot r;
#if defined(__CUDA_ARCH__)
    double sd, cd;
    ::sincos((double)a, &sd, &cd);
    r.s = (float)sd;
    r.c = (float)cd;
#else
    r.s = (float)::sin((double)a);
    r.c = (float)::cos((double)a);
#endif
    return r;
}

MRP_HD V2 rmul(Rot q, V2 v) { return mk(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
MRP_HD V2 rmulT(Rot q, V2 v) { return mk(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
MRP_HD V2 xmul(Xf T, V2 v) {
    float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
    float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
    return mk(x, y);
}
MRP_HD V2 xmulT(Xf T, V2 v) {
    float px = v.x - T.p.x, py = v.y - T.p.y;
    return mk(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
MRP_HD Xf xmulT(Xf A, Xf B) {  // b2MulT(A, B) = inv(A) * B
    Xf C;
    C.q.s = A.q.c * B.q.s - A.q.s * B.q.c;
    C.q.c = A.q.c * B.q.c + A.q.s * B.q.s;
    C.p = rmulT(A.q, B.p - A.p);
    return C;
}

struct Box {  // b2AABB
    float lx, ly, hx, hy;
};
MRP_HD bool overlap(Box a, Box b) {  // b2TestOverlap
    if (b.lx - a.hx > 0.0f || b.ly - a.hy > 0.0f) return false;
    if (a.lx - b.hx > 0.0f || a.ly - b.hy > 0.0f) return false;
    return true;
}
MRP_HD bool contains(Box a, Box b) {  // a.Contains(b)
    return a.lx <= b.lx && a.ly <= b.ly && b.hx <= a.hx && b.hy <= a.hy;
}

// ---------------------------------------------------------------- Philox4x32-10
struct U4 {
    uint32_t x, y, z, w;
};
MRP_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    U4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
enum { kStreamSpawn = 1, kStreamResetAction = 2, kStreamAction = 3 };
// d-th uniform double in [0,1) of the sequence keyed by (seed, stream, global env id, epoch)
MRP_HD double uniform53(uint64_t seed, uint32_t stream, uint64_t env, uint32_t epoch, uint32_t d) {
    U4 r = philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), epoch, d >> 1, (uint32_t)seed,
                         (uint32_t)(seed >> 32) ^ (stream * 0x9E3779B9u));
    uint32_t hi = (d & 1) ? r.z : r.x, lo = (d & 1) ? r.w : r.y;
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

}  // namespace mrp
