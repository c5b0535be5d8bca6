// mrp_collide.cuh — polygon narrowphase (SAT + clip) and dynamic-vs-static time of
// impact for the sm_100a kernels.  Follows Box2D 2.3.x (>= 2.3.1 forks) as specified in
// SURVEY.md Appendix A.5 (b2CollidePolygons) and Appendix E (b2Distance /
// b2SeparationFunction / b2TimeOfImpact) — the arithmetic that pybox2d runs inside
// world.Step (reference mrp00:428, mrp02:478).
//
// Shapes live in a per-CTA shared-memory table (one copy per CTA, read by every lane);
// a shape is 33 floats: [count, vx[8], vy[8], nx[8], ny[8]].
#pragma once
#include "mrp_math.cuh"

namespace mrp {

constexpr int kShapeWords = 33;
MRP_HD int sh_count(const float* sh) { return (int)sh[0]; }
MRP_HD V2 sh_v(const float* sh, int i) { return mk(sh[1 + i], sh[9 + i]); }
MRP_HD V2 sh_n(const float* sh, int i) { return mk(sh[17 + i], sh[25 + i]); }

// contact feature id, 16 bit: indexA | indexB<<4 | typeA<<8 | typeB<<9 (type: 0 vertex, 1 face)
MRP_HD uint32_t idkey(int indexA, int indexB, int typeA, int typeB) {
    return (uint32_t)(indexA | (indexB << 4) | (typeA << 8) | (typeB << 9));
}
MRP_HD uint32_t idkey_flip(uint32_t k) {
    return ((k >> 4) & 15u) | ((k & 15u) << 4) | (((k >> 9) & 1u) << 8) | (((k >> 8) & 1u) << 9);
}

struct Manifold {
    V2 ln, lp;       // localNormal, localPoint
    V2 pt[2];        // points[].localPoint
    uint32_t key[2];
    int pc, type;    // pointCount, 0 = faceA / 1 = faceB
};

struct ClipV {
    V2 v;
    uint32_t key;
};

// b2FindMaxSeparation, brute-force form (A.5)
MRP_HD float find_max_separation(int* edge, const float* sh1, Xf xf1, const float* sh2, Xf xf2) {
    int count1 = sh_count(sh1), count2 = sh_count(sh2);
    Xf xf = xmulT(xf2, xf1);
    int best = 0;
    float maxSep = -FLT_MAX;
    for (int i = 0; i < count1; ++i) {
        V2 n = rmul(xf.q, sh_n(sh1, i));
        V2 v1 = xmul(xf, sh_v(sh1, i));
        float si = FLT_MAX;
        for (int j = 0; j < count2; ++j) {
            float sij = dot(n, sh_v(sh2, j) - v1);
            if (sij < si) si = sij;
        }
        if (si > maxSep) { maxSep = si; best = i; }
    }
    *edge = best;
    return maxSep;
}

MRP_HD int clip_segment(ClipV out[2], const ClipV in[2], V2 normal, float offset, int vertexIndexA) {
    int num = 0;
    float d0 = dot(normal, in[0].v) - offset;
    float d1 = dot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[num++] = in[0];
    if (d1 <= 0.0f) out[num++] = in[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        out[num].v = in[0].v + interp * (in[1].v - in[0].v);
        out[num].key = idkey(vertexIndexA, (int)((in[0].key >> 4) & 15u), 0, 1);
        ++num;
    }
    return num;
}

// b2CollidePolygons
MRP_HDN void collide_polygons(Manifold* m, const float* shA, Xf xfA, const float* shB, Xf xfB) {
    m->pc = 0;
    const float totalRadius = kPolygonRadius + kPolygonRadius;
    int edgeA = 0;
    float sepA = find_max_separation(&edgeA, shA, xfA, shB, xfB);
    if (sepA > totalRadius) return;
    int edgeB = 0;
    float sepB = find_max_separation(&edgeB, shB, xfB, shA, xfA);
    if (sepB > totalRadius) return;

    const float* sh1;
    const float* sh2;
    Xf xf1, xf2;
    int edge1;
    bool flip;
    const float k_tol = 0.1f * kLinearSlop;
    if (sepB > sepA + k_tol) {
        sh1 = shB; sh2 = shA; xf1 = xfB; xf2 = xfA; edge1 = edgeB; m->type = 1; flip = true;
    } else {
        sh1 = shA; sh2 = shB; xf1 = xfA; xf2 = xfB; edge1 = edgeA; m->type = 0; flip = false;
    }
    // b2FindIncidentEdge
    ClipV incident[2];
    {
        int count2 = sh_count(sh2);
        V2 normal1 = rmulT(xf2.q, rmul(xf1.q, sh_n(sh1, edge1)));
        int index = 0;
        float minDot = FLT_MAX;
        for (int i = 0; i < count2; ++i) {
            float d = dot(normal1, sh_n(sh2, i));
            if (d < minDot) { minDot = d; index = i; }
        }
        int i1 = index, i2 = i1 + 1 < count2 ? i1 + 1 : 0;
        incident[0].v = xmul(xf2, sh_v(sh2, i1));
        incident[0].key = idkey(edge1, i1, 1, 0);
        incident[1].v = xmul(xf2, sh_v(sh2, i2));
        incident[1].key = idkey(edge1, i2, 1, 0);
    }
    int count1 = sh_count(sh1);
    int iv1 = edge1, iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
    V2 v11 = sh_v(sh1, iv1), v12 = sh_v(sh1, iv2);
    V2 localTangent = normalized(v12 - v11);
    V2 localNormal = crossVS(localTangent, 1.0f);
    V2 planePoint = 0.5f * (v11 + v12);
    V2 tangent = rmul(xf1.q, localTangent);
    V2 normal = crossVS(tangent, 1.0f);
    v11 = xmul(xf1, v11);
    v12 = xmul(xf1, v12);
    float frontOffset = dot(normal, v11);
    float sideOffset1 = -dot(tangent, v11) + totalRadius;
    float sideOffset2 = dot(tangent, v12) + totalRadius;
    ClipV clip1[2], clip2[2];
    int np = clip_segment(clip1, incident, -tangent, sideOffset1, iv1);
    if (np < 2) return;
    np = clip_segment(clip2, clip1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    m->ln = localNormal;
    m->lp = planePoint;
    int pc = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = dot(normal, clip2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            V2 lp = xmulT(xf2, clip2[i].v);
            uint32_t key = flip ? idkey_flip(clip2[i].key) : clip2[i].key;
            if (pc == 0) { m->pt[0] = lp; m->key[0] = key; } else { m->pt[1] = lp; m->key[1] = key; }
            ++pc;
        }
    }
    m->pc = pc;
}

// ---------------------------------------------------------------- GJK (E.2), useRadii = false
struct Sweep {
    V2 lc, c0, c;
    float a0, a;
};
MRP_HD Xf sweep_xf(const Sweep& s, float beta) {  // b2Sweep::GetTransform
    Xf xf;
    xf.p = (1.0f - beta) * s.c0 + beta * s.c;
    float angle = (1.0f - beta) * s.a0 + beta * s.a;
    // static bodies (the walls: a0 = a = 0) need no sincos: sin(+-0) = +-0, cos(0) = 1 exactly.  Half of the b2Rot::Set
    // evaluations inside b2TimeOfImpact are of this kind (every TOI pair here is dynamic-vs-static).
    if (angle == 0.0f) { xf.q.s = angle; xf.q.c = 1.0f; }
    else xf.q = rot_set(angle);
    V2 r = rmul(xf.q, s.lc);
    xf.p = xf.p - r;
    return xf;
}
MRP_HD int sh_support(const float* sh, V2 d) {  // b2DistanceProxy::GetSupport
    int cnt = sh_count(sh);
    int best = 0;
    float bestValue = dot(sh_v(sh, 0), d);
    for (int i = 1; i < cnt; ++i) {
        float value = dot(sh_v(sh, i), d);
        if (value > bestValue) { best = i; bestValue = value; }
    }
    return best;
}

struct SimplexCache {
    float metric;
    int count;
    int iA[3], iB[3];
};
struct SVert {
    V2 wA, wB, w;
    float a;
    int iA, iB;
};

MRP_HD float simplex_metric(const SVert* v, int count) {
    if (count == 2) return length(v[0].w - v[1].w);
    if (count == 3) return cross(v[1].w - v[0].w, v[2].w - v[0].w);
    return 0.0f;
}

MRP_HDN float gjk_distance(SimplexCache* cache, const float* shA, Xf xfA, const float* shB, Xf xfB) {
    SVert v[3];
    int count = cache->count;
    // ReadCache
    for (int i = 0; i < count; ++i) {
        v[i].iA = cache->iA[i];
        v[i].iB = cache->iB[i];
        v[i].wA = xmul(xfA, sh_v(shA, v[i].iA));
        v[i].wB = xmul(xfB, sh_v(shB, v[i].iB));
        v[i].w = v[i].wB - v[i].wA;
        v[i].a = 0.0f;
    }
    if (count > 1) {
        float metric1 = cache->metric, metric2 = simplex_metric(v, count);
        if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEps) count = 0;
    }
    if (count == 0) {
        v[0].iA = 0; v[0].iB = 0;
        v[0].wA = xmul(xfA, sh_v(shA, 0));
        v[0].wB = xmul(xfB, sh_v(shB, 0));
        v[0].w = v[0].wB - v[0].wA;
        v[0].a = 1.0f;
        count = 1;
    }
    int saveA[3], saveB[3], saveCount = 0;
    int iter = 0;
    while (iter < 20) {
        saveCount = count;
        for (int i = 0; i < saveCount; ++i) { saveA[i] = v[i].iA; saveB[i] = v[i].iB; }
        if (count == 2) {  // Solve2
            V2 w1 = v[0].w, w2 = v[1].w, e12 = w2 - w1;
            float d12_2 = -dot(w1, e12);
            if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; }
            else {
                float d12_1 = dot(w2, e12);
                if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; }
                else {
                    float inv = 1.0f / (d12_1 + d12_2);
                    v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2;
                }
            }
        } else if (count == 3) {  // Solve3
            V2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
            V2 e12 = w2 - w1;
            float d12_1 = dot(w2, e12), d12_2 = -dot(w1, e12);
            V2 e13 = w3 - w1;
            float d13_1 = dot(w3, e13), d13_2 = -dot(w1, e13);
            V2 e23 = w3 - w2;
            float d23_1 = dot(w3, e23), d23_2 = -dot(w2, e23);
            float n123 = cross(e12, e13);
            float d123_1 = n123 * cross(w2, w3);
            float d123_2 = n123 * cross(w3, w1);
            float d123_3 = n123 * cross(w1, w2);
            if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; }
            else if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
                float inv = 1.0f / (d12_1 + d12_2);
                v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2;
            } else if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
                float inv = 1.0f / (d13_1 + d13_2);
                v[0].a = d13_1 * inv; v[2].a = d13_2 * inv; count = 2; v[1] = v[2];
            } else if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; }
            else if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; }
            else if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
                float inv = 1.0f / (d23_1 + d23_2);
                v[1].a = d23_1 * inv; v[2].a = d23_2 * inv; count = 2; v[0] = v[2];
            } else {
                float inv = 1.0f / (d123_1 + d123_2 + d123_3);
                v[0].a = d123_1 * inv; v[1].a = d123_2 * inv; v[2].a = d123_3 * inv; count = 3;
            }
        }
        if (count == 3) break;
        // GetSearchDirection
        V2 d;
        if (count == 1) d = -v[0].w;
        else {
            V2 e12 = v[1].w - v[0].w;
            float sgn = cross(e12, -v[0].w);
            d = sgn > 0.0f ? crossSV(1.0f, e12) : crossVS(e12, 1.0f);
        }
        if (d.x * d.x + d.y * d.y < kEps * kEps) break;
        SVert* nv = v + count;
        nv->iA = sh_support(shA, rmulT(xfA.q, -d));
        nv->wA = xmul(xfA, sh_v(shA, nv->iA));
        nv->iB = sh_support(shB, rmulT(xfB.q, d));
        nv->wB = xmul(xfB, sh_v(shB, nv->iB));
        nv->w = nv->wB - nv->wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i)
            if (nv->iA == saveA[i] && nv->iB == saveB[i]) { duplicate = true; break; }
        if (duplicate) break;
        ++count;
    }
    // GetWitnessPoints
    V2 pA, pB;
    if (count == 1) { pA = v[0].wA; pB = v[0].wB; }
    else if (count == 2) {
        pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
        pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
    } else {
        pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
        pB = pA;
    }
    float distance = length(pA - pB);
    // WriteCache
    cache->metric = simplex_metric(v, count);
    cache->count = count;
    for (int i = 0; i < count; ++i) { cache->iA[i] = v[i].iA; cache->iB[i] = v[i].iB; }
    return distance;
}

// ---------------------------------------------------------------- b2TimeOfImpact (E.3, E.4)
enum { kSepPoints = 0, kSepFaceA = 1, kSepFaceB = 2 };
enum { kToiFailed = 1, kToiOverlapped = 2, kToiTouching = 3, kToiSeparated = 4 };

struct SepFn {
    const float* shA;
    const float* shB;
    Sweep sA, sB;
    V2 localPoint, axis;
    int type;
};

MRP_HD void sep_init(SepFn* f, const SimplexCache* cache, float t1) {
    Xf xfA = sweep_xf(f->sA, t1), xfB = sweep_xf(f->sB, t1);
    if (cache->count == 1) {
        f->type = kSepPoints;
        V2 pointA = xmul(xfA, sh_v(f->shA, cache->iA[0]));
        V2 pointB = xmul(xfB, sh_v(f->shB, cache->iB[0]));
        f->axis = normalized(pointB - pointA);
        f->localPoint = mk(0.0f, 0.0f);
    } else if (cache->iA[0] == cache->iA[1]) {
        f->type = kSepFaceB;
        V2 b1 = sh_v(f->shB, cache->iB[0]), b2 = sh_v(f->shB, cache->iB[1]);
        f->axis = normalized(crossVS(b2 - b1, 1.0f));
        V2 normal = rmul(xfB.q, f->axis);
        f->localPoint = 0.5f * (b1 + b2);
        V2 pointB = xmul(xfB, f->localPoint);
        V2 pointA = xmul(xfA, sh_v(f->shA, cache->iA[0]));
        float s = dot(pointA - pointB, normal);
        if (s < 0.0f) f->axis = -f->axis;
    } else {
        f->type = kSepFaceA;
        V2 a1 = sh_v(f->shA, cache->iA[0]), a2 = sh_v(f->shA, cache->iA[1]);
        f->axis = normalized(crossVS(a2 - a1, 1.0f));
        V2 normal = rmul(xfA.q, f->axis);
        f->localPoint = 0.5f * (a1 + a2);
        V2 pointA = xmul(xfA, f->localPoint);
        V2 pointB = xmul(xfB, sh_v(f->shB, cache->iB[0]));
        float s = dot(pointB - pointA, normal);
        if (s < 0.0f) f->axis = -f->axis;
    }
}
// find == true: FindMinSeparation (support search, writes indices); false: Evaluate
MRP_HD float sep_eval(const SepFn* f, int* indexA, int* indexB, float t, bool find) {
    Xf xfA = sweep_xf(f->sA, t), xfB = sweep_xf(f->sB, t);
    if (f->type == kSepPoints) {
        if (find) {
            *indexA = sh_support(f->shA, rmulT(xfA.q, f->axis));
            *indexB = sh_support(f->shB, rmulT(xfB.q, -f->axis));
        }
        V2 pointA = xmul(xfA, sh_v(f->shA, *indexA)), pointB = xmul(xfB, sh_v(f->shB, *indexB));
        return dot(pointB - pointA, f->axis);
    } else if (f->type == kSepFaceA) {
        V2 normal = rmul(xfA.q, f->axis);
        V2 pointA = xmul(xfA, f->localPoint);
        if (find) {
            *indexA = -1;
            *indexB = sh_support(f->shB, rmulT(xfB.q, -normal));
        }
        V2 pointB = xmul(xfB, sh_v(f->shB, *indexB));
        return dot(pointB - pointA, normal);
    } else {
        V2 normal = rmul(xfB.q, f->axis);
        V2 pointB = xmul(xfB, f->localPoint);
        if (find) {
            *indexB = -1;
            *indexA = sh_support(f->shA, rmulT(xfA.q, -normal));
        }
        V2 pointA = xmul(xfA, sh_v(f->shA, *indexA));
        return dot(pointA - pointB, normal);
    }
}

MRP_HDN int time_of_impact(float* tOut, const float* shA, Sweep sweepA, const float* shB, Sweep sweepB) {
    const float tMax = 1.0f;
    int state = 0;
    *tOut = tMax;
    {  // b2Sweep::Normalize
        const float twoPi = 2.0f * kPi;
        float dA = twoPi * floorf(sweepA.a0 / twoPi);
        sweepA.a0 -= dA; sweepA.a -= dA;
        float dB = twoPi * floorf(sweepB.a0 / twoPi);
        sweepB.a0 -= dB; sweepB.a -= dB;
    }
    const float totalRadius = kPolygonRadius + kPolygonRadius;
    const float target = fmax2(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    const float tolerance = 0.25f * kLinearSlop;
    float t1 = 0.0f;
    int iter = 0;
    SimplexCache cache;
    cache.count = 0;
    cache.metric = 0.0f;
    SepFn fcn;
    fcn.shA = shA; fcn.shB = shB; fcn.sA = sweepA; fcn.sB = sweepB;
    for (;;) {
        Xf xfA = sweep_xf(sweepA, t1), xfB = sweep_xf(sweepB, t1);
        float distance = gjk_distance(&cache, shA, xfA, shB, xfB);
        if (distance <= 0.0f) { state = kToiOverlapped; *tOut = 0.0f; break; }
        if (distance < target + tolerance) { state = kToiTouching; *tOut = t1; break; }
        sep_init(&fcn, &cache, t1);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA = 0, indexB = 0;
            float s2 = sep_eval(&fcn, &indexA, &indexB, t2, true);
            if (s2 > target + tolerance) { state = kToiSeparated; *tOut = tMax; done = true; break; }
            if (s2 > target - tolerance) { t1 = t2; break; }
            float s1 = sep_eval(&fcn, &indexA, &indexB, t1, false);
            if (s1 < target - tolerance) { state = kToiFailed; *tOut = t1; done = true; break; }
            if (s1 <= target + tolerance) { state = kToiTouching; *tOut = t1; done = true; break; }
            int rootIter = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIter & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIter;
                float s = sep_eval(&fcn, &indexA, &indexB, t, false);
                if (fabsf(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                if (rootIter == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == kToiMaxPushBack) break;
        }
        ++iter;
        if (done) break;
        if (iter == 20) { state = kToiFailed; *tOut = t1; break; }
    }
    return state;
}

}  // namespace mrp
