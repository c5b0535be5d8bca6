// oracle/b2core.hpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// CPU restatement of the subset of Box2D 2.3.x (box2d-py, unpinned at
// /root/reference/setup.py:10) that gym_puzzles' MultiRobotPuzzle envs reach through
//   world.Step(1.0/FPS, 6*30, 2*30)            reference mrp00:428, mrp02:478
//   CreateDynamicBody / CreatePolygonFixture   reference mrp00:313-376, mrp02:322-389
//   CreateStaticBody                           reference mrp00:268-274, mrp02:402-410
//   ApplyForce / ApplyTorque / Apply*Impulse   reference mrp00:424, mrp02:454-474
//   contactListener Begin/EndContact           reference mrp00:92-111, mrp02:85-102
// Box2D's sources are NOT in /root/reference and box2d-py is not installable here
// (SURVEY.md §8c), so this follows the published 2.3.x algorithm as specified in
// SURVEY.md Appendix A/E (>=2.3.1 forks: brute-force b2FindMaxSeparation,
// reference-face tolerance 0.1*linearSlop, Pade damping; b2_maxPolygonVertices = 16
// as in pybox2d).  *** PARITY UNPINNED ***: the reference ships no golden vectors for
// this path; this oracle is pinned only by known-answer tests: hand-derived ones
// (tests/test_oracle_kat.py) and ones whose answers come from outside this repository
// (tests/test_b2_external_kats.py: the block solver's LCP solved by hand, a rotating-rod
// time of impact in closed form, and the printed output of the Box2D manual's
// "Hello Box2D" program, which this file reproduces digit for digit).
//
// One deliberate, documented deviation: b2Rot::Set uses correctly-rounded
// sinf/cosf ((float)sin((double)a)); glibc 2.39 sinf differs from that by 1 ulp in
// ~1.3 % of calls (measured), and Box2D itself uses whatever libm the wheel was
// built against, so no single answer is canonical.
//
// Everything is float32, no FMA contraction (build with -ffp-contract=off).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace b2o {

// ---------------------------------------------------------------- settings (A.1)
static const float kPi = 3.14159265359f;
static const float kLinearSlop = 0.005f;
static const float kPolygonRadius = 2.0f * kLinearSlop;
static const float kAabbExtension = 0.1f;
static const float kAabbMultiplier = 2.0f;
static const int kMaxSubSteps = 8;
static const int kMaxTOIContacts = 32;
static const float kVelocityThreshold = 1.0f;
static const float kMaxLinearCorrection = 0.2f;
static const float kMaxTranslation = 2.0f;
static const float kMaxTranslationSquared = kMaxTranslation * kMaxTranslation;
static const float kMaxRotation = 0.5f * kPi;
static const float kMaxRotationSquared = kMaxRotation * kMaxRotation;
static const float kBaumgarte = 0.2f;
static const float kToiBaumgarte = 0.75f;
static const float kEpsilon = FLT_EPSILON;
static const int kMaxPolygonVertices = 16;  // pybox2d's b2Settings.h (upstream: 8)
static const int kMaxVerts = 8;             // storage bound for the shapes used here

// ---------------------------------------------------------------- math (A.0)
struct Vec2 {
    float x, y;
    Vec2() : x(0), y(0) {}
    Vec2(float x_, float y_) : x(x_), y(y_) {}
    float LengthSquared() const { return x * x + y * y; }
    float Length() const { return std::sqrt(x * x + y * y); }
    float Normalize() {
        float len = Length();
        if (len < kEpsilon) return 0.0f;
        float inv = 1.0f / len;
        x *= inv;
        y *= inv;
        return len;
    }
};
inline Vec2 operator+(Vec2 a, Vec2 b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(Vec2 a, Vec2 b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator-(Vec2 a) { return Vec2(-a.x, -a.y); }
inline Vec2 operator*(float s, Vec2 a) { return Vec2(s * a.x, s * a.y); }
inline void operator+=(Vec2& a, Vec2 b) { a.x += b.x; a.y += b.y; }
inline void operator-=(Vec2& a, Vec2 b) { a.x -= b.x; a.y -= b.y; }
inline void operator*=(Vec2& a, float s) { a.x *= s; a.y *= s; }
inline float Dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
inline float Cross(Vec2 a, Vec2 b) { return a.x * b.y - a.y * b.x; }
inline Vec2 Cross(Vec2 a, float s) { return Vec2(s * a.y, -s * a.x); }
inline Vec2 Cross(float s, Vec2 a) { return Vec2(-s * a.y, s * a.x); }
// b2Min / b2Max are ternaries in Box2D (they differ from std::min/max in which signed zero wins)
inline float Min(float a, float b) { return a < b ? a : b; }
inline float Max(float a, float b) { return a > b ? a : b; }
inline Vec2 Min(Vec2 a, Vec2 b) { return Vec2(Min(a.x, b.x), Min(a.y, b.y)); }
inline Vec2 Max(Vec2 a, Vec2 b) { return Vec2(Max(a.x, b.x), Max(a.y, b.y)); }
inline float Clamp(float a, float lo, float hi) { return Max(lo, Min(a, hi)); }
inline float DistanceSquared(Vec2 a, Vec2 b) { Vec2 c = a - b; return Dot(c, c); }
inline float Distance(Vec2 a, Vec2 b) { return (a - b).Length(); }

inline float sin_cr(float a) { return (float)std::sin((double)a); }
inline float cos_cr(float a) { return (float)std::cos((double)a); }

struct Rot {
    float s, c;
    Rot() : s(0), c(1) {}
    void Set(float a) { s = sin_cr(a); c = cos_cr(a); }
};
struct Transform {
    Vec2 p;
    Rot q;
};
inline Vec2 Mul(const Rot& q, Vec2 v) { return Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline Vec2 MulT(const Rot& q, Vec2 v) { return Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
inline Vec2 Mul(const Transform& T, Vec2 v) {
    float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
    float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
    return Vec2(x, y);
}
inline Vec2 MulT(const Transform& T, Vec2 v) {
    float px = v.x - T.p.x, py = v.y - T.p.y;
    return Vec2(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
inline Rot MulT(const Rot& q, const Rot& r) {
    Rot o;
    o.s = q.c * r.s - q.s * r.c;
    o.c = q.c * r.c + q.s * r.s;
    return o;
}
inline Transform MulT(const Transform& A, const Transform& B) {
    Transform C;
    C.q = MulT(A.q, B.q);
    C.p = MulT(A.q, B.p - A.p);
    return C;
}

struct Sweep {
    Vec2 localCenter, c0, c;
    float a0 = 0, a = 0, alpha0 = 0;
    void GetTransform(Transform* xf, float beta) const {  // E.1
        xf->p = (1.0f - beta) * c0 + beta * c;
        float angle = (1.0f - beta) * a0 + beta * a;
        xf->q.Set(angle);
        xf->p -= Mul(xf->q, localCenter);
    }
    void Advance(float alpha) {
        float beta = (alpha - alpha0) / (1.0f - alpha0);
        c0 += beta * (c - c0);
        a0 += beta * (a - a0);
        alpha0 = alpha;
    }
    void Normalize() {
        float twoPi = 2.0f * kPi;
        float d = twoPi * std::floor(a0 / twoPi);
        a0 -= d;
        a -= d;
    }
};

struct AABB {
    Vec2 lo, hi;
    bool Contains(const AABB& o) const {
        bool r = true;
        r = r && lo.x <= o.lo.x;
        r = r && lo.y <= o.lo.y;
        r = r && o.hi.x <= hi.x;
        r = r && o.hi.y <= hi.y;
        return r;
    }
    void Combine(const AABB& a, const AABB& b) {
        lo = Min(a.lo, b.lo);
        hi = Max(a.hi, b.hi);
    }
};
inline bool TestOverlap(const AABB& a, const AABB& b) {
    Vec2 d1 = b.lo - a.hi, d2 = a.lo - b.hi;
    if (d1.x > 0.0f || d1.y > 0.0f) return false;
    if (d2.x > 0.0f || d2.y > 0.0f) return false;
    return true;
}

// ---------------------------------------------------------------- polygon shape (A.2)
struct MassData {
    float mass = 0;
    Vec2 center;
    float I = 0;
};

struct Polygon {
    Vec2 v[kMaxVerts], n[kMaxVerts];
    Vec2 centroid;
    int count = 0;
    float radius = kPolygonRadius;

    void SetAsBox(float hx, float hy) {
        count = 4;
        v[0] = Vec2(-hx, -hy); v[1] = Vec2(hx, -hy); v[2] = Vec2(hx, hy); v[3] = Vec2(-hx, hy);
        n[0] = Vec2(0.0f, -1.0f); n[1] = Vec2(1.0f, 0.0f); n[2] = Vec2(0.0f, 1.0f); n[3] = Vec2(-1.0f, 0.0f);
        centroid = Vec2(0, 0);
    }
    void SetAsBox(float hx, float hy, Vec2 center, float angle) {
        SetAsBox(hx, hy);
        centroid = center;
        Transform xf;
        xf.p = center;
        xf.q.Set(angle);
        for (int i = 0; i < count; ++i) {
            v[i] = Mul(xf, v[i]);
            n[i] = Mul(xf.q, n[i]);
        }
    }
    static Vec2 ComputeCentroid(const Vec2* vs, int cnt) {
        Vec2 c(0, 0);
        float area = 0.0f;
        Vec2 pRef(0, 0);
        const float inv3 = 1.0f / 3.0f;
        for (int i = 0; i < cnt; ++i) {
            Vec2 p1 = pRef, p2 = vs[i], p3 = i + 1 < cnt ? vs[i + 1] : vs[0];
            Vec2 e1 = p2 - p1, e2 = p3 - p1;
            float D = Cross(e1, e2);
            float triangleArea = 0.5f * D;
            area += triangleArea;
            c += triangleArea * inv3 * (p1 + p2 + p3);
        }
        c *= 1.0f / area;
        return c;
    }
    // b2PolygonShape::Set — weld, gift-wrap hull from right-most point
    void Set(const Vec2* pts, int cnt) {
        int nn = std::min(cnt, kMaxVerts);
        Vec2 ps[kMaxVerts];
        int tempCount = 0;
        for (int i = 0; i < nn; ++i) {
            Vec2 p = pts[i];
            bool unique = true;
            for (int j = 0; j < tempCount; ++j)
                if (DistanceSquared(p, ps[j]) < ((0.5f * kLinearSlop) * (0.5f * kLinearSlop))) { unique = false; break; }
            if (unique) ps[tempCount++] = p;
        }
        nn = tempCount;
        int i0 = 0;
        float x0 = ps[0].x;
        for (int i = 1; i < nn; ++i) {
            float x = ps[i].x;
            if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
        }
        int hull[kMaxVerts];
        int m = 0, ih = i0;
        for (;;) {
            hull[m] = ih;
            int ie = 0;
            for (int j = 1; j < nn; ++j) {
                if (ie == ih) { ie = j; continue; }
                Vec2 r = ps[ie] - ps[hull[m]];
                Vec2 w = ps[j] - ps[hull[m]];
                float c = Cross(r, w);
                if (c < 0.0f) ie = j;
                if (c == 0.0f && w.LengthSquared() > r.LengthSquared()) ie = j;
            }
            ++m;
            ih = ie;
            if (ie == i0) break;
        }
        count = m;
        for (int i = 0; i < m; ++i) v[i] = ps[hull[i]];
        for (int i = 0; i < m; ++i) {
            int i2 = i + 1 < m ? i + 1 : 0;
            Vec2 edge = v[i2] - v[i];
            n[i] = Cross(edge, 1.0f);
            n[i].Normalize();
        }
        centroid = ComputeCentroid(v, m);
    }
    void ComputeAABB(AABB* aabb, const Transform& xf) const {
        Vec2 lower = Mul(xf, v[0]), upper = lower;
        for (int i = 1; i < count; ++i) {
            Vec2 p = Mul(xf, v[i]);
            lower = Min(lower, p);
            upper = Max(upper, p);
        }
        Vec2 r(radius, radius);
        aabb->lo = lower - r;
        aabb->hi = upper + r;
    }
    void ComputeMass(MassData* md, float density) const {
        Vec2 center(0, 0);
        float area = 0.0f, I = 0.0f;
        Vec2 s(0, 0);
        for (int i = 0; i < count; ++i) s += v[i];
        s *= 1.0f / count;
        const float k_inv3 = 1.0f / 3.0f;
        for (int i = 0; i < count; ++i) {
            Vec2 e1 = v[i] - s;
            Vec2 e2 = i + 1 < count ? v[i + 1] - s : v[0] - s;
            float D = Cross(e1, e2);
            float triangleArea = 0.5f * D;
            area += triangleArea;
            center += triangleArea * k_inv3 * (e1 + e2);
            float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
            float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
            float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
            I += (0.25f * k_inv3 * D) * (intx2 + inty2);
        }
        md->mass = density * area;
        center *= 1.0f / area;
        md->center = center + s;
        md->I = density * I;
        md->I += md->mass * (Dot(md->center, md->center) - Dot(center, center));
    }
};

// ---------------------------------------------------------------- manifold / collide (A.5)
enum { kFaceA = 0, kFaceB = 1 };
enum { kVertex = 0, kFace = 1 };
struct ContactID {
    uint8_t indexA = 0, indexB = 0, typeA = 0, typeB = 0;
    uint32_t key() const { return indexA | (indexB << 8) | (typeA << 16) | (typeB << 24); }
};
struct ManifoldPoint {
    Vec2 localPoint;
    float normalImpulse = 0, tangentImpulse = 0;
    ContactID id;
};
struct Manifold {
    ManifoldPoint points[2];
    Vec2 localNormal, localPoint;
    int type = kFaceA;
    int pointCount = 0;
};
struct ClipVertex {
    Vec2 v;
    ContactID id;
};

inline int ClipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], Vec2 normal, float offset, int vertexIndexA) {
    int numOut = 0;
    float d0 = Dot(normal, vIn[0].v) - offset;
    float d1 = Dot(normal, vIn[1].v) - offset;
    if (d0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (d1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
        vOut[numOut].id.indexA = (uint8_t)vertexIndexA;
        vOut[numOut].id.indexB = vIn[0].id.indexB;
        vOut[numOut].id.typeA = kVertex;
        vOut[numOut].id.typeB = kFace;
        ++numOut;
    }
    return numOut;
}

// Version fork: 0 = Box2D >= 2.3.1 (what box2d-py >= 2.3.5 tracks; the default and what the CUDA
// kernels implement), 1 = Box2D 2.3.0: hill-climbing b2FindMaxSeparation from the edge facing the other centroid, and the
// reference-face choice `sepB > 0.98 * sepA + 0.001`.  The switch exists to MEASURE what the unpinned choice can cost
// (tests/test_box2d_forks.py): how many env-steps of a rollout change when the other fork is taken.
inline int g_fork_230 = 0;

// Box2D 2.3.0 b2EdgeSeparation: separation of poly2 from edge1 of poly1 along that edge's normal
inline float EdgeSeparation230(const Polygon* p1, const Transform& xf1, int edge1, const Polygon* p2, const Transform& xf2) {
    Vec2 normal1World = Mul(xf1.q, p1->n[edge1]);
    Vec2 normal1 = MulT(xf2.q, normal1World);
    int index = 0;
    float minDot = FLT_MAX;
    for (int i = 0; i < p2->count; ++i) {
        float d = Dot(p2->v[i], normal1);
        if (d < minDot) { minDot = d; index = i; }
    }
    Vec2 v1 = Mul(xf1, p1->v[edge1]);
    Vec2 v2 = Mul(xf2, p2->v[index]);
    return Dot(v2 - v1, normal1World);
}
// Box2D 2.3.0 b2FindMaxSeparation: start at the edge whose normal best faces the other polygon's centroid, then walk
// in the improving direction while the separation strictly improves
inline float FindMaxSeparation230(int* edgeIndex, const Polygon* p1, const Transform& xf1, const Polygon* p2, const Transform& xf2) {
    int count1 = p1->count;
    Vec2 d = Mul(xf2, p2->centroid) - Mul(xf1, p1->centroid);
    Vec2 dLocal1 = MulT(xf1.q, d);
    int edge = 0;
    float maxDot = -FLT_MAX;
    for (int i = 0; i < count1; ++i) {
        float dt = Dot(p1->n[i], dLocal1);
        if (dt > maxDot) { maxDot = dt; edge = i; }
    }
    float s = EdgeSeparation230(p1, xf1, edge, p2, xf2);
    int prevEdge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
    float sPrev = EdgeSeparation230(p1, xf1, prevEdge, p2, xf2);
    int nextEdge = edge + 1 < count1 ? edge + 1 : 0;
    float sNext = EdgeSeparation230(p1, xf1, nextEdge, p2, xf2);
    int bestEdge, increment;
    float bestSeparation;
    if (sPrev > s && sPrev > sNext) { increment = -1; bestEdge = prevEdge; bestSeparation = sPrev; }
    else if (sNext > s) { increment = 1; bestEdge = nextEdge; bestSeparation = sNext; }
    else { *edgeIndex = edge; return s; }
    for (;;) {
        if (increment == -1) edge = bestEdge - 1 >= 0 ? bestEdge - 1 : count1 - 1;
        else edge = bestEdge + 1 < count1 ? bestEdge + 1 : 0;
        s = EdgeSeparation230(p1, xf1, edge, p2, xf2);
        if (s > bestSeparation) { bestEdge = edge; bestSeparation = s; }
        else break;
    }
    *edgeIndex = bestEdge;
    return bestSeparation;
}

// >=2.3.1 brute-force form
inline float FindMaxSeparation(int* edgeIndex, const Polygon* p1, const Transform& xf1, const Polygon* p2, const Transform& xf2) {
    if (g_fork_230) return FindMaxSeparation230(edgeIndex, p1, xf1, p2, xf2);
    int count1 = p1->count, count2 = p2->count;
    Transform xf = MulT(xf2, xf1);
    int bestIndex = 0;
    float maxSeparation = -FLT_MAX;
    for (int i = 0; i < count1; ++i) {
        Vec2 n = Mul(xf.q, p1->n[i]);
        Vec2 v1 = Mul(xf, p1->v[i]);
        float si = FLT_MAX;
        for (int j = 0; j < count2; ++j) {
            float sij = Dot(n, p2->v[j] - v1);
            if (sij < si) si = sij;
        }
        if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
    }
    *edgeIndex = bestIndex;
    return maxSeparation;
}

inline void FindIncidentEdge(ClipVertex c[2], const Polygon* p1, const Transform& xf1, int edge1, const Polygon* p2, const Transform& xf2) {
    int count2 = p2->count;
    Vec2 normal1 = MulT(xf2.q, Mul(xf1.q, p1->n[edge1]));
    int index = 0;
    float minDot = FLT_MAX;
    for (int i = 0; i < count2; ++i) {
        float d = Dot(normal1, p2->n[i]);
        if (d < minDot) { minDot = d; index = i; }
    }
    int i1 = index, i2 = i1 + 1 < count2 ? i1 + 1 : 0;
    c[0].v = Mul(xf2, p2->v[i1]);
    c[0].id.indexA = (uint8_t)edge1; c[0].id.indexB = (uint8_t)i1; c[0].id.typeA = kFace; c[0].id.typeB = kVertex;
    c[1].v = Mul(xf2, p2->v[i2]);
    c[1].id.indexA = (uint8_t)edge1; c[1].id.indexB = (uint8_t)i2; c[1].id.typeA = kFace; c[1].id.typeB = kVertex;
}

inline void CollidePolygons(Manifold* manifold, const Polygon* polyA, const Transform& xfA, const Polygon* polyB, const Transform& xfB) {
    manifold->pointCount = 0;
    float totalRadius = polyA->radius + polyB->radius;
    int edgeA = 0;
    float separationA = FindMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
    if (separationA > totalRadius) return;
    int edgeB = 0;
    float separationB = FindMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
    if (separationB > totalRadius) return;

    const Polygon *poly1, *poly2;
    Transform xf1, xf2;
    int edge1;
    uint8_t flip;
    const float k_tol = 0.1f * kLinearSlop;
    if (g_fork_230 ? separationB > 0.98f * separationA + 0.001f : separationB > separationA + k_tol) {
        poly1 = polyB; poly2 = polyA; xf1 = xfB; xf2 = xfA; edge1 = edgeB;
        manifold->type = kFaceB; flip = 1;
    } else {
        poly1 = polyA; poly2 = polyB; xf1 = xfA; xf2 = xfB; edge1 = edgeA;
        manifold->type = kFaceA; flip = 0;
    }
    ClipVertex incidentEdge[2];
    FindIncidentEdge(incidentEdge, poly1, xf1, edge1, poly2, xf2);
    int count1 = poly1->count;
    int iv1 = edge1, iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
    Vec2 v11 = poly1->v[iv1], v12 = poly1->v[iv2];
    Vec2 localTangent = v12 - v11;
    localTangent.Normalize();
    Vec2 localNormal = Cross(localTangent, 1.0f);
    Vec2 planePoint = 0.5f * (v11 + v12);
    Vec2 tangent = Mul(xf1.q, localTangent);
    Vec2 normal = Cross(tangent, 1.0f);
    v11 = Mul(xf1, v11);
    v12 = Mul(xf1, v12);
    float frontOffset = Dot(normal, v11);
    float sideOffset1 = -Dot(tangent, v11) + totalRadius;
    float sideOffset2 = Dot(tangent, v12) + totalRadius;
    ClipVertex clipPoints1[2], clipPoints2[2];
    int np;
    np = ClipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
    if (np < 2) return;
    np = ClipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    manifold->localNormal = localNormal;
    manifold->localPoint = planePoint;
    int pointCount = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = Dot(normal, clipPoints2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            ManifoldPoint* cp = manifold->points + pointCount;
            cp->localPoint = MulT(xf2, clipPoints2[i].v);
            cp->id = clipPoints2[i].id;
            if (flip) {
                ContactID cf = cp->id;
                cp->id.indexA = cf.indexB; cp->id.indexB = cf.indexA;
                cp->id.typeA = cf.typeB;   cp->id.typeB = cf.typeA;
            }
            ++pointCount;
        }
    }
    manifold->pointCount = pointCount;
}

struct WorldManifold {
    Vec2 normal, points[2];
    void Initialize(const Manifold* m, const Transform& xfA, float radiusA, const Transform& xfB, float radiusB) {
        if (m->pointCount == 0) return;
        if (m->type == kFaceA) {
            normal = Mul(xfA.q, m->localNormal);
            Vec2 planePoint = Mul(xfA, m->localPoint);
            for (int i = 0; i < m->pointCount; ++i) {
                Vec2 clipPoint = Mul(xfB, m->points[i].localPoint);
                Vec2 cA = clipPoint + (radiusA - Dot(clipPoint - planePoint, normal)) * normal;
                Vec2 cB = clipPoint - radiusB * normal;
                points[i] = 0.5f * (cA + cB);
            }
        } else {
            normal = Mul(xfB.q, m->localNormal);
            Vec2 planePoint = Mul(xfB, m->localPoint);
            for (int i = 0; i < m->pointCount; ++i) {
                Vec2 clipPoint = Mul(xfA, m->points[i].localPoint);
                Vec2 cB = clipPoint + (radiusB - Dot(clipPoint - planePoint, normal)) * normal;
                Vec2 cA = clipPoint - radiusA * normal;
                points[i] = 0.5f * (cA + cB);
            }
            normal = -normal;
        }
    }
};

// ---------------------------------------------------------------- GJK distance (E.2)
struct DistanceProxy {
    const Vec2* verts = nullptr;
    int count = 0;
    float radius = 0;
    void Set(const Polygon* p) { verts = p->v; count = p->count; radius = p->radius; }
    int GetSupport(Vec2 d) const {
        int best = 0;
        float bestValue = Dot(verts[0], d);
        for (int i = 1; i < count; ++i) {
            float value = Dot(verts[i], d);
            if (value > bestValue) { best = i; bestValue = value; }
        }
        return best;
    }
    Vec2 GetVertex(int i) const { return verts[i]; }
};
struct SimplexCache {
    float metric = 0;
    uint16_t count = 0;
    uint8_t indexA[3] = {0, 0, 0}, indexB[3] = {0, 0, 0};
};
struct SimplexVertex {
    Vec2 wA, wB, w;
    float a = 0;
    int indexA = 0, indexB = 0;
};
struct Simplex {
    SimplexVertex v[3];
    int count = 0;
    float GetMetric() const {
        switch (count) {
            case 1: return 0.0f;
            case 2: return Distance(v[0].w, v[1].w);
            case 3: return Cross(v[1].w - v[0].w, v[2].w - v[0].w);
            default: return 0.0f;
        }
    }
    void ReadCache(const SimplexCache* cache, const DistanceProxy* pA, const Transform& xfA, const DistanceProxy* pB, const Transform& xfB) {
        count = cache->count;
        for (int i = 0; i < count; ++i) {
            SimplexVertex* s = v + i;
            s->indexA = cache->indexA[i];
            s->indexB = cache->indexB[i];
            s->wA = Mul(xfA, pA->GetVertex(s->indexA));
            s->wB = Mul(xfB, pB->GetVertex(s->indexB));
            s->w = s->wB - s->wA;
            s->a = 0.0f;
        }
        if (count > 1) {
            float metric1 = cache->metric, metric2 = GetMetric();
            if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
        }
        if (count == 0) {
            SimplexVertex* s = v;
            s->indexA = 0; s->indexB = 0;
            s->wA = Mul(xfA, pA->GetVertex(0));
            s->wB = Mul(xfB, pB->GetVertex(0));
            s->w = s->wB - s->wA;
            s->a = 1.0f;
            count = 1;
        }
    }
    void WriteCache(SimplexCache* cache) const {
        cache->metric = GetMetric();
        cache->count = (uint16_t)count;
        for (int i = 0; i < count; ++i) { cache->indexA[i] = (uint8_t)v[i].indexA; cache->indexB[i] = (uint8_t)v[i].indexB; }
    }
    Vec2 GetSearchDirection() const {
        if (count == 1) return -v[0].w;
        Vec2 e12 = v[1].w - v[0].w;
        float sgn = Cross(e12, -v[0].w);
        if (sgn > 0.0f) return Cross(1.0f, e12);
        return Cross(e12, 1.0f);
    }
    void GetWitnessPoints(Vec2* pA, Vec2* pB) const {
        switch (count) {
            case 1: *pA = v[0].wA; *pB = v[0].wB; break;
            case 2:
                *pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
                *pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
                break;
            case 3:
                *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
                *pB = *pA;
                break;
        }
    }
    void Solve2() {
        Vec2 w1 = v[0].w, w2 = v[1].w, e12 = w2 - w1;
        float d12_2 = -Dot(w1, e12);
        if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        float d12_1 = Dot(w2, e12);
        if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        float inv = 1.0f / (d12_1 + d12_2);
        v[0].a = d12_1 * inv;
        v[1].a = d12_2 * inv;
        count = 2;
    }
    void Solve3() {
        Vec2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
        Vec2 e12 = w2 - w1;
        float w1e12 = Dot(w1, e12), w2e12 = Dot(w2, e12);
        float d12_1 = w2e12, d12_2 = -w1e12;
        Vec2 e13 = w3 - w1;
        float w1e13 = Dot(w1, e13), w3e13 = Dot(w3, e13);
        float d13_1 = w3e13, d13_2 = -w1e13;
        Vec2 e23 = w3 - w2;
        float w2e23 = Dot(w2, e23), w3e23 = Dot(w3, e23);
        float d23_1 = w3e23, d23_2 = -w2e23;
        float n123 = Cross(e12, e13);
        float d123_1 = n123 * Cross(w2, w3);
        float d123_2 = n123 * Cross(w3, w1);
        float d123_3 = n123 * Cross(w1, w2);
        if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
            float inv = 1.0f / (d12_1 + d12_2);
            v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2; return;
        }
        if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
            float inv = 1.0f / (d13_1 + d13_2);
            v[0].a = d13_1 * inv; v[2].a = d13_2 * inv; count = 2; v[1] = v[2]; return;
        }
        if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; return; }
        if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
            float inv = 1.0f / (d23_1 + d23_2);
            v[1].a = d23_1 * inv; v[2].a = d23_2 * inv; count = 2; v[0] = v[2]; return;
        }
        float inv = 1.0f / (d123_1 + d123_2 + d123_3);
        v[0].a = d123_1 * inv; v[1].a = d123_2 * inv; v[2].a = d123_3 * inv; count = 3;
    }
};

// b2Distance with useRadii=false (the only mode b2TimeOfImpact uses)
inline float DistanceGJK(SimplexCache* cache, const DistanceProxy* pA, const Transform& xfA, const DistanceProxy* pB, const Transform& xfB) {
    Simplex simplex;
    simplex.ReadCache(cache, pA, xfA, pB, xfB);
    SimplexVertex* vertices = simplex.v;
    const int k_maxIters = 20;
    int saveA[3], saveB[3], saveCount = 0;
    int iter = 0;
    while (iter < k_maxIters) {
        saveCount = simplex.count;
        for (int i = 0; i < saveCount; ++i) { saveA[i] = vertices[i].indexA; saveB[i] = vertices[i].indexB; }
        switch (simplex.count) {
            case 1: break;
            case 2: simplex.Solve2(); break;
            case 3: simplex.Solve3(); break;
        }
        if (simplex.count == 3) break;
        Vec2 d = simplex.GetSearchDirection();
        if (d.LengthSquared() < kEpsilon * kEpsilon) break;
        SimplexVertex* vertex = vertices + simplex.count;
        vertex->indexA = pA->GetSupport(MulT(xfA.q, -d));
        vertex->wA = Mul(xfA, pA->GetVertex(vertex->indexA));
        vertex->indexB = pB->GetSupport(MulT(xfB.q, d));
        vertex->wB = Mul(xfB, pB->GetVertex(vertex->indexB));
        vertex->w = vertex->wB - vertex->wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i)
            if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) { duplicate = true; break; }
        if (duplicate) break;
        ++simplex.count;
    }
    Vec2 pointA, pointB;
    simplex.GetWitnessPoints(&pointA, &pointB);
    float distance = Distance(pointA, pointB);
    simplex.WriteCache(cache);
    return distance;
}

// ---------------------------------------------------------------- time of impact (E.3, E.4)
struct SeparationFunction {
    enum Type { kPoints, kFaceAType, kFaceBType };
    const DistanceProxy *pA, *pB;
    Sweep sweepA, sweepB;
    Type type;
    Vec2 localPoint, axis;

    void Initialize(const SimplexCache* cache, const DistanceProxy* proxyA, const Sweep& sA, const DistanceProxy* proxyB, const Sweep& sB, float t1) {
        pA = proxyA; pB = proxyB;
        int count = cache->count;
        sweepA = sA; sweepB = sB;
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t1);
        sweepB.GetTransform(&xfB, t1);
        if (count == 1) {
            type = kPoints;
            Vec2 pointA = Mul(xfA, pA->GetVertex(cache->indexA[0]));
            Vec2 pointB = Mul(xfB, pB->GetVertex(cache->indexB[0]));
            axis = pointB - pointA;
            axis.Normalize();
        } else if (cache->indexA[0] == cache->indexA[1]) {
            type = kFaceBType;
            Vec2 b1 = pB->GetVertex(cache->indexB[0]), b2 = pB->GetVertex(cache->indexB[1]);
            axis = Cross(b2 - b1, 1.0f);
            axis.Normalize();
            Vec2 normal = Mul(xfB.q, axis);
            localPoint = 0.5f * (b1 + b2);
            Vec2 pointB = Mul(xfB, localPoint);
            Vec2 pointA = Mul(xfA, pA->GetVertex(cache->indexA[0]));
            float s = Dot(pointA - pointB, normal);
            if (s < 0.0f) axis = -axis;
        } else {
            type = kFaceAType;
            Vec2 a1 = pA->GetVertex(cache->indexA[0]), a2 = pA->GetVertex(cache->indexA[1]);
            axis = Cross(a2 - a1, 1.0f);
            axis.Normalize();
            Vec2 normal = Mul(xfA.q, axis);
            localPoint = 0.5f * (a1 + a2);
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 pointB = Mul(xfB, pB->GetVertex(cache->indexB[0]));
            float s = Dot(pointB - pointA, normal);
            if (s < 0.0f) axis = -axis;
        }
    }
    float FindMinSeparation(int* indexA, int* indexB, float t) const {
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t);
        sweepB.GetTransform(&xfB, t);
        switch (type) {
            case kPoints: {
                Vec2 axisA = MulT(xfA.q, axis), axisB = MulT(xfB.q, -axis);
                *indexA = pA->GetSupport(axisA);
                *indexB = pB->GetSupport(axisB);
                Vec2 pointA = Mul(xfA, pA->GetVertex(*indexA)), pointB = Mul(xfB, pB->GetVertex(*indexB));
                return Dot(pointB - pointA, axis);
            }
            case kFaceAType: {
                Vec2 normal = Mul(xfA.q, axis);
                Vec2 pointA = Mul(xfA, localPoint);
                Vec2 axisB = MulT(xfB.q, -normal);
                *indexA = -1;
                *indexB = pB->GetSupport(axisB);
                Vec2 pointB = Mul(xfB, pB->GetVertex(*indexB));
                return Dot(pointB - pointA, normal);
            }
            default: {
                Vec2 normal = Mul(xfB.q, axis);
                Vec2 pointB = Mul(xfB, localPoint);
                Vec2 axisA = MulT(xfA.q, -normal);
                *indexB = -1;
                *indexA = pA->GetSupport(axisA);
                Vec2 pointA = Mul(xfA, pA->GetVertex(*indexA));
                return Dot(pointA - pointB, normal);
            }
        }
    }
    float Evaluate(int indexA, int indexB, float t) const {
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t);
        sweepB.GetTransform(&xfB, t);
        switch (type) {
            case kPoints: {
                Vec2 pointA = Mul(xfA, pA->GetVertex(indexA)), pointB = Mul(xfB, pB->GetVertex(indexB));
                return Dot(pointB - pointA, axis);
            }
            case kFaceAType: {
                Vec2 normal = Mul(xfA.q, axis);
                Vec2 pointA = Mul(xfA, localPoint);
                Vec2 pointB = Mul(xfB, pB->GetVertex(indexB));
                return Dot(pointB - pointA, normal);
            }
            default: {
                Vec2 normal = Mul(xfB.q, axis);
                Vec2 pointB = Mul(xfB, localPoint);
                Vec2 pointA = Mul(xfA, pA->GetVertex(indexA));
                return Dot(pointA - pointB, normal);
            }
        }
    }
};

enum TOIState { kToiUnknown, kToiFailed, kToiOverlapped, kToiTouching, kToiSeparated };

inline TOIState TimeOfImpact(float* tOut, const DistanceProxy* proxyA, const DistanceProxy* proxyB, Sweep sweepA, Sweep sweepB, float tMax) {
    TOIState state = kToiUnknown;
    *tOut = tMax;
    sweepA.Normalize();
    sweepB.Normalize();
    float totalRadius = proxyA->radius + proxyB->radius;
    float target = Max(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    float tolerance = 0.25f * kLinearSlop;
    float t1 = 0.0f;
    const int k_maxIterations = 20;
    int iter = 0;
    SimplexCache cache;
    cache.count = 0;
    for (;;) {
        Transform xfA, xfB;
        sweepA.GetTransform(&xfA, t1);
        sweepB.GetTransform(&xfB, t1);
        float distance = DistanceGJK(&cache, proxyA, xfA, proxyB, xfB);
        if (distance <= 0.0f) { state = kToiOverlapped; *tOut = 0.0f; break; }
        if (distance < target + tolerance) { state = kToiTouching; *tOut = t1; break; }
        SeparationFunction fcn;
        fcn.Initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = fcn.FindMinSeparation(&indexA, &indexB, t2);
            if (s2 > target + tolerance) { state = kToiSeparated; *tOut = tMax; done = true; break; }
            if (s2 > target - tolerance) { t1 = t2; break; }
            float s1 = fcn.Evaluate(indexA, indexB, t1);
            if (s1 < target - tolerance) { state = kToiFailed; *tOut = t1; done = true; break; }
            if (s1 <= target + tolerance) { state = kToiTouching; *tOut = t1; done = true; break; }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIterCount;
                float s = fcn.Evaluate(indexA, indexB, t);
                if (std::fabs(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == kMaxPolygonVertices) break;
        }
        ++iter;
        if (done) break;
        if (iter == k_maxIterations) { state = kToiFailed; *tOut = t1; break; }
    }
    return state;
}

// ---------------------------------------------------------------- bodies / fixtures / contacts
enum BodyType { kStatic = 0, kDynamic = 2 };

struct Contact;
struct Fixture {
    int body = -1;
    Polygon shape;
    float density = 0, friction = 0.2f, restitution = 0;
    AABB aabb;  // proxy->aabb (tight/swept)
};
struct Body {
    BodyType type = kStatic;
    Transform xf;
    Sweep sweep;
    Vec2 v, force;
    float w = 0, torque = 0;
    float mass = 0, invMass = 0, I = 0, invI = 0;
    float linearDamping = 0, angularDamping = 0;
    bool islandFlag = false;
    int islandIndex = 0;
    std::vector<int> fixtures;      // newest first (b2Body::m_fixtureList)
    std::vector<Contact*> contacts; // contact edges, newest first (b2Body::m_contactList)

    void SynchronizeTransform() {
        xf.q.Set(sweep.a);
        xf.p = sweep.c - Mul(xf.q, sweep.localCenter);
    }
    void Advance(float alpha) {
        sweep.Advance(alpha);
        sweep.c = sweep.c0;
        sweep.a = sweep.a0;
        xf.q.Set(sweep.a);
        xf.p = sweep.c - Mul(xf.q, sweep.localCenter);
    }
    Vec2 GetWorldPoint(Vec2 lp) const { return Mul(xf, lp); }
    Vec2 GetWorldVector(Vec2 lv) const { return Mul(xf.q, lv); }
    float GetInertia() const { return I + mass * Dot(sweep.localCenter, sweep.localCenter); }
    void ApplyForce(Vec2 f, Vec2 point) {
        if (type != kDynamic) return;
        force += f;
        torque += Cross(point - sweep.c, f);
    }
    void ApplyTorque(float t) { if (type == kDynamic) torque += t; }
    void ApplyLinearImpulse(Vec2 imp, Vec2 point) {
        if (type != kDynamic) return;
        v += invMass * imp;
        w += invI * Cross(point - sweep.c, imp);
    }
    void ApplyAngularImpulse(float imp) { if (type == kDynamic) w += invI * imp; }
};

struct Contact {
    int fA, fB;  // fixture indices (fA = lower proxy id)
    int bA, bB;
    bool touching = false, enabled = true, islandFlag = false, toiFlag = false;
    int toiCount = 0;
    float toi = 1.0f;
    Manifold manifold;
    float friction = 0, restitution = 0;
};

struct ContactListener {
    virtual ~ContactListener() {}
    virtual void BeginContact(Contact*) {}
    virtual void EndContact(Contact*) {}
};

struct TimeStep {
    float dt, inv_dt, dtRatio;
    int velocityIterations, positionIterations;
    bool warmStarting;
};

// ---------------------------------------------------------------- contact solver (A.8)
struct VelocityConstraintPoint {
    Vec2 rA, rB;
    float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias;
};
struct VelocityConstraint {
    VelocityConstraintPoint points[2];
    Vec2 normal;
    float nm_exx, nm_exy, nm_eyx, nm_eyy;  // normalMass (Mat22: ex=(exx,exy), ey=(eyx,eyy))
    float K_exx, K_exy, K_eyx, K_eyy;
    int indexA, indexB;
    float invMassA, invMassB, invIA, invIB, friction, restitution;
    int pointCount, contactIndex;
};
struct PositionConstraint {
    Vec2 localPoints[2], localNormal, localPoint;
    int indexA, indexB;
    float invMassA, invMassB;
    Vec2 localCenterA, localCenterB;
    float invIA, invIB;
    int type;
    float radiusA, radiusB;
    int pointCount;
};
struct Position { Vec2 c; float a; };
struct Velocity { Vec2 v; float w; };

struct World;

struct ContactSolver {
    TimeStep step;
    std::vector<Position>* positions;
    std::vector<Velocity>* velocities;
    std::vector<Contact*>* contacts;
    std::vector<VelocityConstraint> vcs;
    std::vector<PositionConstraint> pcs;
    const World* world;

    ContactSolver(const TimeStep& st, std::vector<Contact*>* cs, std::vector<Position>* ps, std::vector<Velocity>* vs, const World* w);
    void InitializeVelocityConstraints();
    void WarmStart();
    void SolveVelocityConstraints();
    void StoreImpulses();
    bool SolvePositionConstraints();
    bool SolveTOIPositionConstraints(int toiIndexA, int toiIndexB);
};

// ---------------------------------------------------------------- world (A.3, A.4, A.6, A.7, A.10)
struct World {
    std::vector<Body> bodies;       // creation order; b2World::m_bodyList order is the reverse
    std::vector<Fixture> fixtures;  // creation order == proxy id order in a fresh world (A.3)
    std::vector<AABB> fat;          // fat AABB per proxy
    std::vector<int> moveBuffer;
    std::vector<Contact*> contactList;  // head first
    ContactListener* listener = nullptr;
    bool newFixture = false;
    float inv_dt0 = 0.0f;
    bool warmStarting = true, continuousPhysics = true;
    // instrumentation for tests
    long stat_toi_events = 0, stat_toi_calls = 0, stat_pos_iters = 0;
    // workload probes (used to size the CUDA design; do not alter results)
    bool probe = false;
    long stat_islands = 0, stat_islands_c[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // histogram by #contacts (7 = 7+)
    long stat_period[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // [p] = islands whose velocity sweeps enter a cycle of period p <= 8; [0]: 9..32
    // position solver: islands, islands that use all iterations, and of those: state cycle of period 1 / 2 / 3..8 found, sum of
    // the iteration at which it was found; iterations spent by all islands
    long stat_pos_islands = 0, stat_pos_full = 0, stat_pos_cyc[3] = {0, 0, 0}, stat_pos_cyc_at = 0, stat_pos_iter_sum = 0;
    long stat_vel_cyc_at = 0;   // sum over islands with a velocity cycle of period >= 2 of the sweep that closed it
    long stat_fix_iters = 0, stat_fix_never = 0, stat_contact_sum = 0, stat_touch_sum = 0, stat_steps = 0;

    ~World() { for (Contact* c : contactList) delete c; }
    World() {}
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    int CreateBody(BodyType type, Vec2 position, float angle, float linDamp, float angDamp) {
        Body b;
        b.type = type;
        b.xf.p = position;
        b.xf.q.Set(angle);
        b.sweep.localCenter = Vec2(0, 0);
        b.sweep.c0 = b.sweep.c = position;
        b.sweep.a0 = b.sweep.a = angle;
        b.sweep.alpha0 = 0;
        b.linearDamping = linDamp;
        b.angularDamping = angDamp;
        if (type == kDynamic) { b.mass = 1.0f; b.invMass = 1.0f; }
        bodies.push_back(b);
        return (int)bodies.size() - 1;
    }
    int CreateFixture(int body, const Polygon& shape, float density, float friction, float restitution) {
        Fixture f;
        f.body = body;
        f.shape = shape;
        f.density = density;
        f.friction = friction;
        f.restitution = restitution;
        f.shape.ComputeAABB(&f.aabb, bodies[body].xf);
        fixtures.push_back(f);
        int id = (int)fixtures.size() - 1;
        AABB fa;
        Vec2 r(kAabbExtension, kAabbExtension);
        fa.lo = f.aabb.lo - r;
        fa.hi = f.aabb.hi + r;
        fat.push_back(fa);
        moveBuffer.push_back(id);
        bodies[body].fixtures.insert(bodies[body].fixtures.begin(), id);
        if (density > 0.0f) ResetMassData(body);
        newFixture = true;
        return id;
    }
    void ResetMassData(int bi) {
        Body& b = bodies[bi];
        b.mass = 0; b.invMass = 0; b.I = 0; b.invI = 0;
        b.sweep.localCenter = Vec2(0, 0);
        if (b.type == kStatic) {
            b.sweep.c0 = b.xf.p; b.sweep.c = b.xf.p; b.sweep.a0 = b.sweep.a;
            return;
        }
        Vec2 localCenter(0, 0);
        for (int fi : b.fixtures) {
            const Fixture& f = fixtures[fi];
            if (f.density == 0.0f) continue;
            MassData md;
            f.shape.ComputeMass(&md, f.density);
            b.mass += md.mass;
            localCenter += md.mass * md.center;
            b.I += md.I;
        }
        if (b.mass > 0.0f) {
            b.invMass = 1.0f / b.mass;
            localCenter *= b.invMass;
        } else {
            b.mass = 1.0f; b.invMass = 1.0f;
        }
        if (b.I > 0.0f) {
            b.I -= b.mass * Dot(localCenter, localCenter);
            b.invI = 1.0f / b.I;
        } else {
            b.I = 0; b.invI = 0;
        }
        Vec2 oldCenter = b.sweep.c;
        b.sweep.localCenter = localCenter;
        b.sweep.c0 = b.sweep.c = Mul(b.xf, b.sweep.localCenter);
        b.v += Cross(b.w, b.sweep.c - oldCenter);
    }

    // ---- contact manager
    void DestroyContact(Contact* c) {
        if (listener && c->touching) listener->EndContact(c);
        contactList.erase(std::find(contactList.begin(), contactList.end(), c));
        auto& ea = bodies[c->bA].contacts;
        ea.erase(std::find(ea.begin(), ea.end(), c));
        auto& eb = bodies[c->bB].contacts;
        eb.erase(std::find(eb.begin(), eb.end(), c));
        delete c;
    }
    void AddPair(int proxyA, int proxyB) {
        int bodyA = fixtures[proxyA].body, bodyB = fixtures[proxyB].body;
        if (bodyA == bodyB) return;
        for (Contact* e : bodies[bodyB].contacts) {
            int other = (e->bA == bodyB) ? e->bB : e->bA;
            if (other == bodyA) {
                if (e->fA == proxyA && e->fB == proxyB) return;
                if (e->fA == proxyB && e->fB == proxyA) return;
            }
        }
        if (bodies[bodyA].type != kDynamic && bodies[bodyB].type != kDynamic) return;
        Contact* c = new Contact();
        c->fA = proxyA; c->fB = proxyB; c->bA = bodyA; c->bB = bodyB;
        c->friction = std::sqrt(fixtures[proxyA].friction * fixtures[proxyB].friction);
        c->restitution = Max(fixtures[proxyA].restitution, fixtures[proxyB].restitution);
        c->manifold.pointCount = 0;
        contactList.insert(contactList.begin(), c);
        bodies[bodyA].contacts.insert(bodies[bodyA].contacts.begin(), c);
        bodies[bodyB].contacts.insert(bodies[bodyB].contacts.begin(), c);
    }
    void FindNewContacts() {
        std::vector<std::pair<int, int>> pairs;
        int np = (int)fixtures.size();
        for (int q : moveBuffer) {
            for (int p = 0; p < np; ++p) {
                if (p == q) continue;
                if (TestOverlap(fat[p], fat[q])) pairs.push_back({std::min(p, q), std::max(p, q)});
            }
        }
        moveBuffer.clear();
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        for (auto& pr : pairs) AddPair(pr.first, pr.second);
    }
    void UpdateContact(Contact* c) {  // b2Contact::Update (A.4)
        Manifold old = c->manifold;
        c->enabled = true;
        bool wasTouching = c->touching;
        const Fixture &fa = fixtures[c->fA], &fb = fixtures[c->fB];
        CollidePolygons(&c->manifold, &fa.shape, bodies[c->bA].xf, &fb.shape, bodies[c->bB].xf);
        bool touching = c->manifold.pointCount > 0;
        for (int i = 0; i < c->manifold.pointCount; ++i) {
            ManifoldPoint* mp2 = c->manifold.points + i;
            mp2->normalImpulse = 0.0f;
            mp2->tangentImpulse = 0.0f;
            for (int j = 0; j < old.pointCount; ++j) {
                const ManifoldPoint* mp1 = old.points + j;
                if (mp1->id.key() == mp2->id.key()) {
                    mp2->normalImpulse = mp1->normalImpulse;
                    mp2->tangentImpulse = mp1->tangentImpulse;
                    break;
                }
            }
        }
        c->touching = touching;
        if (!wasTouching && touching && listener) listener->BeginContact(c);
        if (wasTouching && !touching && listener) listener->EndContact(c);
    }
    void Collide() {
        size_t i = 0;
        while (i < contactList.size()) {
            Contact* c = contactList[i];
            bool activeA = bodies[c->bA].type != kStatic, activeB = bodies[c->bB].type != kStatic;
            if (!activeA && !activeB) { ++i; continue; }
            if (!TestOverlap(fat[c->fA], fat[c->fB])) {
                DestroyContact(c);  // erases index i; next contact slides into i
                continue;
            }
            UpdateContact(c);
            ++i;
        }
    }
    void SynchronizeFixtures(int bi) {
        Body& b = bodies[bi];
        Transform xf1;
        xf1.q.Set(b.sweep.a0);
        xf1.p = b.sweep.c0 - Mul(xf1.q, b.sweep.localCenter);
        for (int fi : b.fixtures) {
            Fixture& f = fixtures[fi];
            AABB a1, a2;
            f.shape.ComputeAABB(&a1, xf1);
            f.shape.ComputeAABB(&a2, b.xf);
            f.aabb.Combine(a1, a2);
            Vec2 displacement = b.xf.p - xf1.p;
            // b2DynamicTree::MoveProxy
            if (fat[fi].Contains(f.aabb)) continue;
            AABB nb = f.aabb;
            Vec2 r(kAabbExtension, kAabbExtension);
            nb.lo = nb.lo - r;
            nb.hi = nb.hi + r;
            Vec2 d = kAabbMultiplier * displacement;
            if (d.x < 0.0f) nb.lo.x += d.x; else nb.hi.x += d.x;
            if (d.y < 0.0f) nb.lo.y += d.y; else nb.hi.y += d.y;
            fat[fi] = nb;
            moveBuffer.push_back(fi);
        }
    }

    // Harness entry (not a Box2D function): put the dynamic bodies, their proxies' fat AABBs and the contact list into
    // a given between-steps state.  Bodies [0, nDyn) and fixtures [0, nDynFix) are the dynamic ones (creation order:
    // the env builders create them before the walls).  bodies6 = (c.x, c.y, a, v.x, v.y, w) per body; contact records
    // are include/mrp_state.h's 14 words, list head first.  The move buffer is empty between two steps.
    void LoadState(int nDyn, const float* bodies6, int nDynFix, const float* fat4, int nc, const uint32_t* cwords) {
        for (Contact* c : contactList) delete c;
        contactList.clear();
        for (Body& b : bodies) b.contacts.clear();
        for (int b = 0; b < nDyn; ++b) {
            Body& B = bodies[b];
            B.sweep.c = Vec2(bodies6[6 * b + 0], bodies6[6 * b + 1]);
            B.sweep.a = bodies6[6 * b + 2];
            B.sweep.c0 = B.sweep.c; B.sweep.a0 = B.sweep.a;
            B.v = Vec2(bodies6[6 * b + 3], bodies6[6 * b + 4]);
            B.w = bodies6[6 * b + 5];
            B.force = Vec2(0, 0); B.torque = 0.0f;
            B.SynchronizeTransform();
        }
        for (int f = 0; f < nDynFix; ++f) {
            fat[f].lo = Vec2(fat4[4 * f + 0], fat4[4 * f + 1]);
            fat[f].hi = Vec2(fat4[4 * f + 2], fat4[4 * f + 3]);
        }
        moveBuffer.clear();
        newFixture = false;
        inv_dt0 = 50.0f;
        // record k is list position k (head first): rebuild tail -> head
        for (int k = nc - 1; k >= 0; --k) {
            const uint32_t* cw = cwords + 14 * k;
            const float* cf = (const float*)cw;
            Contact* c = new Contact();
            c->fA = cw[0] & 0xff; c->fB = (cw[0] >> 8) & 0xff;
            c->bA = fixtures[c->fA].body; c->bB = fixtures[c->fB].body;
            c->touching = ((cw[0] >> 16) & 1) != 0;
            c->manifold.type = (cw[0] >> 17) & 1;
            c->manifold.pointCount = (cw[0] >> 18) & 3;
            c->friction = std::sqrt(fixtures[c->fA].friction * fixtures[c->fB].friction);
            c->restitution = 0.0f;
            c->manifold.localNormal = Vec2(cf[2], cf[3]);
            c->manifold.localPoint = Vec2(cf[4], cf[5]);
            for (int j = 0; j < 2; ++j) {
                uint32_t k16 = (cw[1] >> (16 * j)) & 0xffff;
                ManifoldPoint& p = c->manifold.points[j];
                p.id.indexA = k16 & 15; p.id.indexB = (k16 >> 4) & 15; p.id.typeA = (k16 >> 8) & 1; p.id.typeB = (k16 >> 9) & 1;
                p.localPoint = Vec2(cf[6 + 4 * j], cf[7 + 4 * j]);
                p.normalImpulse = cf[8 + 4 * j];
                p.tangentImpulse = cf[9 + 4 * j];
            }
            contactList.insert(contactList.begin(), c);
            bodies[c->bA].contacts.insert(bodies[c->bA].contacts.begin(), c);
            bodies[c->bB].contacts.insert(bodies[c->bB].contacts.begin(), c);
        }
    }

    void Step(float dt, int velocityIterations, int positionIterations);
    void Solve(const TimeStep& step);
    void SolveTOI(const TimeStep& step);
    void SolveIsland(std::vector<int>& ibodies, std::vector<Contact*>& icontacts, const TimeStep& step);
    void SolveTOIIsland(std::vector<int>& ibodies, std::vector<Contact*>& icontacts, const TimeStep& subStep, int toiIndexA, int toiIndexB);
    void ClearForces() {
        for (Body& b : bodies) { b.force = Vec2(0, 0); b.torque = 0.0f; }
    }
};

// ---------------------------------------------------------------- contact solver impl
inline ContactSolver::ContactSolver(const TimeStep& st, std::vector<Contact*>* cs, std::vector<Position>* ps, std::vector<Velocity>* vs, const World* w)
    : step(st), positions(ps), velocities(vs), contacts(cs), world(w) {
    int count = (int)cs->size();
    vcs.resize(count);
    pcs.resize(count);
    for (int i = 0; i < count; ++i) {
        Contact* contact = (*cs)[i];
        const Fixture &fA = w->fixtures[contact->fA], &fB = w->fixtures[contact->fB];
        const Body &bodyA = w->bodies[contact->bA], &bodyB = w->bodies[contact->bB];
        const Manifold* manifold = &contact->manifold;
        int pointCount = manifold->pointCount;
        VelocityConstraint* vc = &vcs[i];
        vc->friction = contact->friction;
        vc->restitution = contact->restitution;
        vc->indexA = bodyA.islandIndex;
        vc->indexB = bodyB.islandIndex;
        vc->invMassA = bodyA.invMass; vc->invMassB = bodyB.invMass;
        vc->invIA = bodyA.invI; vc->invIB = bodyB.invI;
        vc->contactIndex = i;
        vc->pointCount = pointCount;
        vc->K_exx = vc->K_exy = vc->K_eyx = vc->K_eyy = 0;
        vc->nm_exx = vc->nm_exy = vc->nm_eyx = vc->nm_eyy = 0;
        PositionConstraint* pc = &pcs[i];
        pc->indexA = bodyA.islandIndex; pc->indexB = bodyB.islandIndex;
        pc->invMassA = bodyA.invMass; pc->invMassB = bodyB.invMass;
        pc->localCenterA = bodyA.sweep.localCenter; pc->localCenterB = bodyB.sweep.localCenter;
        pc->invIA = bodyA.invI; pc->invIB = bodyB.invI;
        pc->localNormal = manifold->localNormal;
        pc->localPoint = manifold->localPoint;
        pc->pointCount = pointCount;
        pc->radiusA = fA.shape.radius; pc->radiusB = fB.shape.radius;
        pc->type = manifold->type;
        for (int j = 0; j < pointCount; ++j) {
            const ManifoldPoint* cp = manifold->points + j;
            VelocityConstraintPoint* vcp = vc->points + j;
            if (step.warmStarting) {
                vcp->normalImpulse = step.dtRatio * cp->normalImpulse;
                vcp->tangentImpulse = step.dtRatio * cp->tangentImpulse;
            } else {
                vcp->normalImpulse = 0.0f;
                vcp->tangentImpulse = 0.0f;
            }
            vcp->rA = Vec2(0, 0); vcp->rB = Vec2(0, 0);
            vcp->normalMass = 0; vcp->tangentMass = 0; vcp->velocityBias = 0;
            pc->localPoints[j] = cp->localPoint;
        }
    }
}

inline void ContactSolver::InitializeVelocityConstraints() {
    for (size_t i = 0; i < vcs.size(); ++i) {
        VelocityConstraint* vc = &vcs[i];
        PositionConstraint* pc = &pcs[i];
        float radiusA = pc->radiusA, radiusB = pc->radiusB;
        const Manifold* manifold = &(*contacts)[vc->contactIndex]->manifold;
        int indexA = vc->indexA, indexB = vc->indexB;
        float mA = vc->invMassA, mB = vc->invMassB, iA = vc->invIA, iB = vc->invIB;
        Vec2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
        Vec2 cA = (*positions)[indexA].c; float aA = (*positions)[indexA].a;
        Vec2 vA = (*velocities)[indexA].v; float wA = (*velocities)[indexA].w;
        Vec2 cB = (*positions)[indexB].c; float aB = (*positions)[indexB].a;
        Vec2 vB = (*velocities)[indexB].v; float wB = (*velocities)[indexB].w;
        Transform xfA, xfB;
        xfA.q.Set(aA); xfB.q.Set(aB);
        xfA.p = cA - Mul(xfA.q, localCenterA);
        xfB.p = cB - Mul(xfB.q, localCenterB);
        WorldManifold wm;
        wm.Initialize(manifold, xfA, radiusA, xfB, radiusB);
        vc->normal = wm.normal;
        int pointCount = vc->pointCount;
        for (int j = 0; j < pointCount; ++j) {
            VelocityConstraintPoint* vcp = vc->points + j;
            vcp->rA = wm.points[j] - cA;
            vcp->rB = wm.points[j] - cB;
            float rnA = Cross(vcp->rA, vc->normal), rnB = Cross(vcp->rB, vc->normal);
            float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
            vcp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
            Vec2 tangent = Cross(vc->normal, 1.0f);
            float rtA = Cross(vcp->rA, tangent), rtB = Cross(vcp->rB, tangent);
            float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
            vcp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
            vcp->velocityBias = 0.0f;
            float vRel = Dot(vc->normal, vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA));
            if (vRel < -kVelocityThreshold) vcp->velocityBias = -vc->restitution * vRel;
        }
        if (vc->pointCount == 2) {
            VelocityConstraintPoint *vcp1 = vc->points + 0, *vcp2 = vc->points + 1;
            float rn1A = Cross(vcp1->rA, vc->normal), rn1B = Cross(vcp1->rB, vc->normal);
            float rn2A = Cross(vcp2->rA, vc->normal), rn2B = Cross(vcp2->rB, vc->normal);
            float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
            float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
            float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
            const float k_maxConditionNumber = 1000.0f;
            if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
                vc->K_exx = k11; vc->K_exy = k12; vc->K_eyx = k12; vc->K_eyy = k22;
                // b2Mat22::GetInverse
                float a = k11, b = k12, c = k12, d = k22;
                float det = a * d - b * c;
                if (det != 0.0f) det = 1.0f / det;
                vc->nm_exx = det * d; vc->nm_eyx = -det * b;
                vc->nm_exy = -det * c; vc->nm_eyy = det * a;
            } else {
                vc->pointCount = 1;
            }
        }
    }
}

inline void ContactSolver::WarmStart() {
    for (size_t i = 0; i < vcs.size(); ++i) {
        VelocityConstraint* vc = &vcs[i];
        int indexA = vc->indexA, indexB = vc->indexB;
        float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
        int pointCount = vc->pointCount;
        Vec2 vA = (*velocities)[indexA].v; float wA = (*velocities)[indexA].w;
        Vec2 vB = (*velocities)[indexB].v; float wB = (*velocities)[indexB].w;
        Vec2 normal = vc->normal;
        Vec2 tangent = Cross(normal, 1.0f);
        for (int j = 0; j < pointCount; ++j) {
            VelocityConstraintPoint* vcp = vc->points + j;
            Vec2 P = vcp->normalImpulse * normal + vcp->tangentImpulse * tangent;
            wA -= iA * Cross(vcp->rA, P);
            vA -= mA * P;
            wB += iB * Cross(vcp->rB, P);
            vB += mB * P;
        }
        (*velocities)[indexA].v = vA; (*velocities)[indexA].w = wA;
        (*velocities)[indexB].v = vB; (*velocities)[indexB].w = wB;
    }
}

inline void ContactSolver::SolveVelocityConstraints() {
    for (size_t i = 0; i < vcs.size(); ++i) {
        VelocityConstraint* vc = &vcs[i];
        int indexA = vc->indexA, indexB = vc->indexB;
        float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
        int pointCount = vc->pointCount;
        Vec2 vA = (*velocities)[indexA].v; float wA = (*velocities)[indexA].w;
        Vec2 vB = (*velocities)[indexB].v; float wB = (*velocities)[indexB].w;
        Vec2 normal = vc->normal;
        Vec2 tangent = Cross(normal, 1.0f);
        float friction = vc->friction;
        for (int j = 0; j < pointCount; ++j) {
            VelocityConstraintPoint* vcp = vc->points + j;
            Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
            float vt = Dot(dv, tangent) - 0.0f;  // tangentSpeed = 0
            float lambda = vcp->tangentMass * (-vt);
            float maxFriction = friction * vcp->normalImpulse;
            float newImpulse = Clamp(vcp->tangentImpulse + lambda, -maxFriction, maxFriction);
            lambda = newImpulse - vcp->tangentImpulse;
            vcp->tangentImpulse = newImpulse;
            Vec2 P = lambda * tangent;
            vA -= mA * P;
            wA -= iA * Cross(vcp->rA, P);
            vB += mB * P;
            wB += iB * Cross(vcp->rB, P);
        }
        if (pointCount == 1) {
            for (int j = 0; j < pointCount; ++j) {
                VelocityConstraintPoint* vcp = vc->points + j;
                Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
                float vn = Dot(dv, normal);
                float lambda = -vcp->normalMass * (vn - vcp->velocityBias);
                float newImpulse = Max(vcp->normalImpulse + lambda, 0.0f);
                lambda = newImpulse - vcp->normalImpulse;
                vcp->normalImpulse = newImpulse;
                Vec2 P = lambda * normal;
                vA -= mA * P;
                wA -= iA * Cross(vcp->rA, P);
                vB += mB * P;
                wB += iB * Cross(vcp->rB, P);
            }
        } else {
            VelocityConstraintPoint *cp1 = vc->points + 0, *cp2 = vc->points + 1;
            Vec2 a(cp1->normalImpulse, cp2->normalImpulse);
            Vec2 dv1 = vB + Cross(wB, cp1->rB) - vA - Cross(wA, cp1->rA);
            Vec2 dv2 = vB + Cross(wB, cp2->rB) - vA - Cross(wA, cp2->rA);
            float vn1 = Dot(dv1, normal), vn2 = Dot(dv2, normal);
            Vec2 b;
            b.x = vn1 - cp1->velocityBias;
            b.y = vn2 - cp2->velocityBias;
            // b -= b2Mul(K, a)
            b.x -= vc->K_exx * a.x + vc->K_eyx * a.y;
            b.y -= vc->K_exy * a.x + vc->K_eyy * a.y;
            auto apply = [&](Vec2 x) {
                Vec2 d = x - a;
                Vec2 P1 = d.x * normal, P2 = d.y * normal;
                vA -= mA * (P1 + P2);
                wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
                vB += mB * (P1 + P2);
                wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
                cp1->normalImpulse = x.x;
                cp2->normalImpulse = x.y;
            };
            for (;;) {
                // case 1: x = -normalMass * b
                Vec2 x;
                x.x = -(vc->nm_exx * b.x + vc->nm_eyx * b.y);
                x.y = -(vc->nm_exy * b.x + vc->nm_eyy * b.y);
                if (x.x >= 0.0f && x.y >= 0.0f) { apply(x); break; }
                // case 2
                x.x = -cp1->normalMass * b.x;
                x.y = 0.0f;
                vn1 = 0.0f;
                vn2 = vc->K_exy * x.x + b.y;
                if (x.x >= 0.0f && vn2 >= 0.0f) { apply(x); break; }
                // case 3
                x.x = 0.0f;
                x.y = -cp2->normalMass * b.y;
                vn1 = vc->K_eyx * x.y + b.x;
                vn2 = 0.0f;
                if (x.y >= 0.0f && vn1 >= 0.0f) { apply(x); break; }
                // case 4
                x.x = 0.0f; x.y = 0.0f;
                vn1 = b.x; vn2 = b.y;
                if (vn1 >= 0.0f && vn2 >= 0.0f) { apply(x); break; }
                break;
            }
        }
        (*velocities)[indexA].v = vA; (*velocities)[indexA].w = wA;
        (*velocities)[indexB].v = vB; (*velocities)[indexB].w = wB;
    }
}

inline void ContactSolver::StoreImpulses() {
    for (size_t i = 0; i < vcs.size(); ++i) {
        VelocityConstraint* vc = &vcs[i];
        Manifold* manifold = &(*contacts)[vc->contactIndex]->manifold;
        for (int j = 0; j < vc->pointCount; ++j) {
            manifold->points[j].normalImpulse = vc->points[j].normalImpulse;
            manifold->points[j].tangentImpulse = vc->points[j].tangentImpulse;
        }
    }
}

struct PositionSolverManifold {
    Vec2 normal, point;
    float separation;
    void Initialize(const PositionConstraint* pc, const Transform& xfA, const Transform& xfB, int index) {
        if (pc->type == kFaceA) {
            normal = Mul(xfA.q, pc->localNormal);
            Vec2 planePoint = Mul(xfA, pc->localPoint);
            Vec2 clipPoint = Mul(xfB, pc->localPoints[index]);
            separation = Dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
            point = clipPoint;
        } else {
            normal = Mul(xfB.q, pc->localNormal);
            Vec2 planePoint = Mul(xfB, pc->localPoint);
            Vec2 clipPoint = Mul(xfA, pc->localPoints[index]);
            separation = Dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
            point = clipPoint;
            normal = -normal;
        }
    }
};

inline bool ContactSolver::SolvePositionConstraints() {
    float minSeparation = 0.0f;
    for (size_t i = 0; i < pcs.size(); ++i) {
        PositionConstraint* pc = &pcs[i];
        int indexA = pc->indexA, indexB = pc->indexB;
        Vec2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
        float mA = pc->invMassA, iA = pc->invIA, mB = pc->invMassB, iB = pc->invIB;
        int pointCount = pc->pointCount;
        Vec2 cA = (*positions)[indexA].c; float aA = (*positions)[indexA].a;
        Vec2 cB = (*positions)[indexB].c; float aB = (*positions)[indexB].a;
        for (int j = 0; j < pointCount; ++j) {
            Transform xfA, xfB;
            xfA.q.Set(aA); xfB.q.Set(aB);
            xfA.p = cA - Mul(xfA.q, localCenterA);
            xfB.p = cB - Mul(xfB.q, localCenterB);
            PositionSolverManifold psm;
            psm.Initialize(pc, xfA, xfB, j);
            Vec2 normal = psm.normal, point = psm.point;
            float separation = psm.separation;
            Vec2 rA = point - cA, rB = point - cB;
            minSeparation = Min(minSeparation, separation);
            float C = Clamp(kBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
            float rnA = Cross(rA, normal), rnB = Cross(rB, normal);
            float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
            float impulse = K > 0.0f ? -C / K : 0.0f;
            Vec2 P = impulse * normal;
            cA -= mA * P;
            aA -= iA * Cross(rA, P);
            cB += mB * P;
            aB += iB * Cross(rB, P);
        }
        (*positions)[indexA].c = cA; (*positions)[indexA].a = aA;
        (*positions)[indexB].c = cB; (*positions)[indexB].a = aB;
    }
    return minSeparation >= -3.0f * kLinearSlop;
}

inline bool ContactSolver::SolveTOIPositionConstraints(int toiIndexA, int toiIndexB) {
    float minSeparation = 0.0f;
    for (size_t i = 0; i < pcs.size(); ++i) {
        PositionConstraint* pc = &pcs[i];
        int indexA = pc->indexA, indexB = pc->indexB;
        Vec2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
        int pointCount = pc->pointCount;
        float mA = 0.0f, iA = 0.0f;
        if (indexA == toiIndexA || indexA == toiIndexB) { mA = pc->invMassA; iA = pc->invIA; }
        float mB = 0.0f, iB = 0.0f;
        if (indexB == toiIndexA || indexB == toiIndexB) { mB = pc->invMassB; iB = pc->invIB; }
        Vec2 cA = (*positions)[indexA].c; float aA = (*positions)[indexA].a;
        Vec2 cB = (*positions)[indexB].c; float aB = (*positions)[indexB].a;
        for (int j = 0; j < pointCount; ++j) {
            Transform xfA, xfB;
            xfA.q.Set(aA); xfB.q.Set(aB);
            xfA.p = cA - Mul(xfA.q, localCenterA);
            xfB.p = cB - Mul(xfB.q, localCenterB);
            PositionSolverManifold psm;
            psm.Initialize(pc, xfA, xfB, j);
            Vec2 normal = psm.normal, point = psm.point;
            float separation = psm.separation;
            Vec2 rA = point - cA, rB = point - cB;
            minSeparation = Min(minSeparation, separation);
            float C = Clamp(kToiBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
            float rnA = Cross(rA, normal), rnB = Cross(rB, normal);
            float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
            float impulse = K > 0.0f ? -C / K : 0.0f;
            Vec2 P = impulse * normal;
            cA -= mA * P;
            aA -= iA * Cross(rA, P);
            cB += mB * P;
            aB += iB * Cross(rB, P);
        }
        (*positions)[indexA].c = cA; (*positions)[indexA].a = aA;
        (*positions)[indexB].c = cB; (*positions)[indexB].a = aB;
    }
    return minSeparation >= -1.5f * kLinearSlop;
}

// ---------------------------------------------------------------- islands
inline void World::SolveIsland(std::vector<int>& ibodies, std::vector<Contact*>& icontacts, const TimeStep& step) {
    float h = step.dt;
    int nb = (int)ibodies.size();
    std::vector<Position> positions(nb);
    std::vector<Velocity> velocities(nb);
    for (int i = 0; i < nb; ++i) {
        Body& b = bodies[ibodies[i]];
        Vec2 c = b.sweep.c; float a = b.sweep.a;
        Vec2 v = b.v; float w = b.w;
        b.sweep.c0 = b.sweep.c;
        b.sweep.a0 = b.sweep.a;
        if (b.type == kDynamic) {
            // gravity = (0,0), gravityScale = 1
            v += h * (1.0f * Vec2(0.0f, 0.0f) + b.invMass * b.force);
            w += h * b.invI * b.torque;
            v *= 1.0f / (1.0f + h * b.linearDamping);
            w *= 1.0f / (1.0f + h * b.angularDamping);
        }
        positions[i].c = c; positions[i].a = a;
        velocities[i].v = v; velocities[i].w = w;
    }
    ContactSolver solver(step, &icontacts, &positions, &velocities, this);
    solver.InitializeVelocityConstraints();
    if (step.warmStarting) solver.WarmStart();
    if (probe && !icontacts.empty()) {
        ++stat_islands;
        ++stat_islands_c[std::min<size_t>(icontacts.size(), 7)];
        ContactSolver s2 = solver;
        std::vector<Velocity> v2 = velocities;
        s2.velocities = &v2;
        int fixed = -1;
        std::vector<std::vector<char>> hist;
        auto snap = [&]() {
            std::vector<char> b(sizeof(Velocity) * v2.size() + sizeof(VelocityConstraint) * s2.vcs.size());
            std::memcpy(b.data(), v2.data(), sizeof(Velocity) * v2.size());
            std::memcpy(b.data() + sizeof(Velocity) * v2.size(), s2.vcs.data(), sizeof(VelocityConstraint) * s2.vcs.size());
            return b;
        };
        hist.push_back(snap());
        for (int i = 0; i < step.velocityIterations && fixed < 0; ++i) {
            s2.SolveVelocityConstraints();
            std::vector<char> cur = snap();
            for (int p = 1; p <= 32 && p <= (int)hist.size(); ++p)
                if (hist[hist.size() - p] == cur) { fixed = i + 1; ++stat_period[p <= 8 ? p : 0]; if (p > 1) stat_vel_cyc_at += i + 1; break; }
            hist.push_back(cur);
        }
        if (fixed < 0) { ++stat_fix_never; stat_fix_iters += step.velocityIterations; } else stat_fix_iters += fixed;
    }
    for (int i = 0; i < step.velocityIterations; ++i) solver.SolveVelocityConstraints();
    solver.StoreImpulses();
    for (int i = 0; i < nb; ++i) {
        Vec2 c = positions[i].c; float a = positions[i].a;
        Vec2 v = velocities[i].v; float w = velocities[i].w;
        Vec2 translation = h * v;
        if (Dot(translation, translation) > kMaxTranslationSquared) {
            float ratio = kMaxTranslation / translation.Length();
            v *= ratio;
        }
        float rotation = h * w;
        if (rotation * rotation > kMaxRotationSquared) {
            float ratio = kMaxRotation / std::fabs(rotation);
            w *= ratio;
        }
        c += h * v;
        a += h * w;
        positions[i].c = c; positions[i].a = a;
        velocities[i].v = v; velocities[i].w = w;
    }
    if (probe && !icontacts.empty()) {   // would the position iterations enter a bitwise cycle before they run out?
        std::vector<Position> p2 = positions;
        ContactSolver s2 = solver;
        s2.positions = &p2;
        std::vector<std::vector<char>> hist;
        auto snap = [&]() { std::vector<char> b(sizeof(Position) * p2.size()); std::memcpy(b.data(), p2.data(), b.size()); return b; };
        hist.push_back(snap());
        int used = 0, cyc_p = 0, cyc_at = 0;
        bool okay = false;
        for (int i = 0; i < step.positionIterations; ++i) {
            ++used;
            okay = s2.SolvePositionConstraints();
            if (okay) break;
            std::vector<char> cur = snap();
            if (!cyc_p)
                for (int p = 1; p <= 8 && p <= (int)hist.size(); ++p)
                    if (hist[hist.size() - p] == cur) { cyc_p = p; cyc_at = i + 1; break; }
            hist.push_back(cur);
        }
        ++stat_pos_islands;
        stat_pos_iter_sum += used;
        if (!okay) {
            ++stat_pos_full;
            if (cyc_p) { ++stat_pos_cyc[cyc_p == 1 ? 0 : (cyc_p == 2 ? 1 : 2)]; stat_pos_cyc_at += cyc_at; }
        }
    }
    for (int i = 0; i < step.positionIterations; ++i) {
        ++stat_pos_iters;
        bool contactsOkay = solver.SolvePositionConstraints();
        if (contactsOkay) break;
    }
    for (int i = 0; i < nb; ++i) {
        Body& b = bodies[ibodies[i]];
        b.sweep.c = positions[i].c;
        b.sweep.a = positions[i].a;
        b.v = velocities[i].v;
        b.w = velocities[i].w;
        b.SynchronizeTransform();
    }
}

inline void World::SolveTOIIsland(std::vector<int>& ibodies, std::vector<Contact*>& icontacts, const TimeStep& subStep, int toiIndexA, int toiIndexB) {
    int nb = (int)ibodies.size();
    std::vector<Position> positions(nb);
    std::vector<Velocity> velocities(nb);
    for (int i = 0; i < nb; ++i) {
        Body& b = bodies[ibodies[i]];
        positions[i].c = b.sweep.c; positions[i].a = b.sweep.a;
        velocities[i].v = b.v; velocities[i].w = b.w;
    }
    ContactSolver solver(subStep, &icontacts, &positions, &velocities, this);
    for (int i = 0; i < subStep.positionIterations; ++i) {
        bool contactsOkay = solver.SolveTOIPositionConstraints(toiIndexA, toiIndexB);
        if (contactsOkay) break;
    }
    bodies[ibodies[toiIndexA]].sweep.c0 = positions[toiIndexA].c;
    bodies[ibodies[toiIndexA]].sweep.a0 = positions[toiIndexA].a;
    bodies[ibodies[toiIndexB]].sweep.c0 = positions[toiIndexB].c;
    bodies[ibodies[toiIndexB]].sweep.a0 = positions[toiIndexB].a;
    solver.InitializeVelocityConstraints();
    for (int i = 0; i < subStep.velocityIterations; ++i) solver.SolveVelocityConstraints();
    float h = subStep.dt;
    for (int i = 0; i < nb; ++i) {
        Vec2 c = positions[i].c; float a = positions[i].a;
        Vec2 v = velocities[i].v; float w = velocities[i].w;
        Vec2 translation = h * v;
        if (Dot(translation, translation) > kMaxTranslationSquared) {
            float ratio = kMaxTranslation / translation.Length();
            v *= ratio;
        }
        float rotation = h * w;
        if (rotation * rotation > kMaxRotationSquared) {
            float ratio = kMaxRotation / std::fabs(rotation);
            w *= ratio;
        }
        c += h * v;
        a += h * w;
        Body& b = bodies[ibodies[i]];
        b.sweep.c = c; b.sweep.a = a;
        b.v = v; b.w = w;
        b.SynchronizeTransform();
    }
}

inline void World::Solve(const TimeStep& step) {
    for (Body& b : bodies) b.islandFlag = false;
    for (Contact* c : contactList) c->islandFlag = false;
    std::vector<int> stack, ibodies;
    std::vector<Contact*> icontacts;
    // b2World::m_bodyList is newest first
    for (int seed = (int)bodies.size() - 1; seed >= 0; --seed) {
        Body& sb = bodies[seed];
        if (sb.islandFlag) continue;
        if (sb.type == kStatic) continue;
        ibodies.clear();
        icontacts.clear();
        stack.clear();
        stack.push_back(seed);
        sb.islandFlag = true;
        while (!stack.empty()) {
            int bi = stack.back();
            stack.pop_back();
            Body& b = bodies[bi];
            b.islandIndex = (int)ibodies.size();
            ibodies.push_back(bi);
            if (b.type == kStatic) continue;
            for (Contact* contact : b.contacts) {
                if (contact->islandFlag) continue;
                if (!contact->enabled || !contact->touching) continue;
                icontacts.push_back(contact);
                contact->islandFlag = true;
                int other = (contact->bA == bi) ? contact->bB : contact->bA;
                if (bodies[other].islandFlag) continue;
                stack.push_back(other);
                bodies[other].islandFlag = true;
            }
        }
        SolveIsland(ibodies, icontacts, step);
        for (int bi : ibodies)
            if (bodies[bi].type == kStatic) bodies[bi].islandFlag = false;
    }
    for (int bi = (int)bodies.size() - 1; bi >= 0; --bi) {
        if (!bodies[bi].islandFlag) continue;
        if (bodies[bi].type == kStatic) continue;
        SynchronizeFixtures(bi);
    }
    FindNewContacts();
}

inline void World::SolveTOI(const TimeStep& step) {
    for (Body& b : bodies) { b.islandFlag = false; b.sweep.alpha0 = 0.0f; }
    for (Contact* c : contactList) { c->toiFlag = false; c->islandFlag = false; c->toiCount = 0; c->toi = 1.0f; }
    for (;;) {
        Contact* minContact = nullptr;
        float minAlpha = 1.0f;
        for (Contact* c : contactList) {
            if (!c->enabled) continue;
            if (c->toiCount > kMaxSubSteps) continue;
            float alpha = 1.0f;
            if (c->toiFlag) {
                alpha = c->toi;
            } else {
                Body &bA = bodies[c->bA], &bB = bodies[c->bB];
                bool activeA = bA.type != kStatic, activeB = bB.type != kStatic;
                if (!activeA && !activeB) continue;
                bool collideA = bA.type != kDynamic, collideB = bB.type != kDynamic;  // no bullets
                if (!collideA && !collideB) continue;
                float alpha0 = bA.sweep.alpha0;
                if (bA.sweep.alpha0 < bB.sweep.alpha0) {
                    alpha0 = bB.sweep.alpha0;
                    bA.sweep.Advance(alpha0);
                } else if (bB.sweep.alpha0 < bA.sweep.alpha0) {
                    alpha0 = bA.sweep.alpha0;
                    bB.sweep.Advance(alpha0);
                }
                DistanceProxy pA, pB;
                pA.Set(&fixtures[c->fA].shape);
                pB.Set(&fixtures[c->fB].shape);
                float t;
                ++stat_toi_calls;
                TOIState st = TimeOfImpact(&t, &pA, &pB, bA.sweep, bB.sweep, 1.0f);
                float beta = t;
                if (st == kToiTouching) alpha = Min(alpha0 + (1.0f - alpha0) * beta, 1.0f);
                else alpha = 1.0f;
                c->toi = alpha;
                c->toiFlag = true;
            }
            if (alpha < minAlpha) { minContact = c; minAlpha = alpha; }
        }
        if (minContact == nullptr || 1.0f - 10.0f * kEpsilon < minAlpha) break;
        ++stat_toi_events;
        int ibA = minContact->bA, ibB = minContact->bB;
        Body &bA = bodies[ibA], &bB = bodies[ibB];
        Sweep backup1 = bA.sweep, backup2 = bB.sweep;
        bA.Advance(minAlpha);
        bB.Advance(minAlpha);
        UpdateContact(minContact);
        minContact->toiFlag = false;
        ++minContact->toiCount;
        if (!minContact->enabled || !minContact->touching) {
            minContact->enabled = false;
            bA.sweep = backup1;
            bB.sweep = backup2;
            bA.SynchronizeTransform();
            bB.SynchronizeTransform();
            continue;
        }
        std::vector<int> ibodies;
        std::vector<Contact*> icontacts;
        bA.islandIndex = 0; ibodies.push_back(ibA);
        bB.islandIndex = 1; ibodies.push_back(ibB);
        icontacts.push_back(minContact);
        bA.islandFlag = true; bB.islandFlag = true; minContact->islandFlag = true;
        int two[2] = {ibA, ibB};
        for (int k = 0; k < 2; ++k) {
            Body& body = bodies[two[k]];
            if (body.type != kDynamic) continue;
            // iterate over a copy: Update never edits edge lists, but stay safe
            std::vector<Contact*> edges = body.contacts;
            for (Contact* contact : edges) {
                if ((int)ibodies.size() == 2 * kMaxTOIContacts) break;
                if ((int)icontacts.size() == kMaxTOIContacts) break;
                if (contact->islandFlag) continue;
                int oi = (contact->bA == two[k]) ? contact->bB : contact->bA;
                Body& other = bodies[oi];
                if (other.type == kDynamic) continue;  // no bullets
                Sweep backup = other.sweep;
                if (!other.islandFlag) other.Advance(minAlpha);
                UpdateContact(contact);
                if (!contact->enabled) { other.sweep = backup; other.SynchronizeTransform(); continue; }
                if (!contact->touching) { other.sweep = backup; other.SynchronizeTransform(); continue; }
                contact->islandFlag = true;
                icontacts.push_back(contact);
                if (other.islandFlag) continue;
                other.islandFlag = true;
                other.islandIndex = (int)ibodies.size();
                ibodies.push_back(oi);
            }
        }
        TimeStep subStep;
        subStep.dt = (1.0f - minAlpha) * step.dt;
        subStep.inv_dt = 1.0f / subStep.dt;
        subStep.dtRatio = 1.0f;
        subStep.positionIterations = 20;
        subStep.velocityIterations = step.velocityIterations;
        subStep.warmStarting = false;
        SolveTOIIsland(ibodies, icontacts, subStep, bA.islandIndex, bB.islandIndex);
        for (int bi : ibodies) {
            Body& body = bodies[bi];
            body.islandFlag = false;
            if (body.type != kDynamic) continue;
            SynchronizeFixtures(bi);
            for (Contact* ce : body.contacts) { ce->toiFlag = false; ce->islandFlag = false; }
        }
        FindNewContacts();
    }
}

inline void World::Step(float dt, int velocityIterations, int positionIterations) {
    if (newFixture) {
        FindNewContacts();
        newFixture = false;
    }
    TimeStep step;
    step.dt = dt;
    step.velocityIterations = velocityIterations;
    step.positionIterations = positionIterations;
    step.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
    step.dtRatio = inv_dt0 * dt;
    step.warmStarting = warmStarting;
    Collide();
    if (probe) {
        ++stat_steps;
        stat_contact_sum += (long)contactList.size();
        for (Contact* c : contactList) stat_touch_sum += c->touching ? 1 : 0;
    }
    if (step.dt > 0.0f) Solve(step);
    if (continuousPhysics && step.dt > 0.0f) SolveTOI(step);
    if (step.dt > 0.0f) inv_dt0 = step.inv_dt;
    ClearForces();
}

}  // namespace b2o
