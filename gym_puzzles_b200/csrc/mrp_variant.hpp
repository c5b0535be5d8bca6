// mrp_variant.hpp — host-side construction of the per-variant constants the kernels
// use: polygon shapes, mass data, fixture/wall tables.  Float32 throughout, same
// operation order as Box2D 2.3.x b2PolygonShape::{Set, SetAsBox, ComputeMass} and
// b2Body::ResetMassData, applied to the shapes the reference builds in
// _generate_blocks / _generate_agents / _generate_boundary
// (reference mrp00:260-275,299-378; mrp02:313-411).
#pragma once
#include <math.h>
#include <string.h>

#include "mrp_sim.cuh"

namespace mrp {

struct HostPoly {
    float vx[8], vy[8], nx[8], ny[8];
    int count;
};

inline void poly_box(HostPoly* p, float hx, float hy, float cx, float cy) {
    // SetAsBox(hx, hy, center, angle = 0): v = Mul(xf, v) with q = (sin 0, cos 0) = (0, 1)
    const float bx[4] = {-hx, hx, hx, -hx}, by[4] = {-hy, -hy, hy, hy};
    const float nx[4] = {0.0f, 1.0f, 0.0f, -1.0f}, ny[4] = {-1.0f, 0.0f, 1.0f, 0.0f};
    p->count = 4;
    for (int i = 0; i < 8; ++i) { p->vx[i] = p->vy[i] = p->nx[i] = p->ny[i] = 0.0f; }
    const float s = 0.0f, c = 1.0f;
    for (int i = 0; i < 4; ++i) {
        p->vx[i] = (c * bx[i] - s * by[i]) + cx;
        p->vy[i] = (s * bx[i] + c * by[i]) + cy;
        p->nx[i] = c * nx[i] - s * ny[i];
        p->ny[i] = s * nx[i] + c * ny[i];
    }
}
inline void poly_box_plain(HostPoly* p, float hx, float hy) {  // SetAsBox(hx, hy)
    const float bx[4] = {-hx, hx, hx, -hx}, by[4] = {-hy, -hy, hy, hy};
    const float nx[4] = {0.0f, 1.0f, 0.0f, -1.0f}, ny[4] = {-1.0f, 0.0f, 1.0f, 0.0f};
    p->count = 4;
    for (int i = 0; i < 8; ++i) { p->vx[i] = p->vy[i] = p->nx[i] = p->ny[i] = 0.0f; }
    for (int i = 0; i < 4; ++i) { p->vx[i] = bx[i]; p->vy[i] = by[i]; p->nx[i] = nx[i]; p->ny[i] = ny[i]; }
}
// b2PolygonShape::Set: gift wrapping from the right-most (lowest on ties) point, CCW
inline void poly_hull(HostPoly* p, const float* px, const float* py, int n) {
    int i0 = 0;
    for (int i = 1; i < n; ++i)
        if (px[i] > px[i0] || (px[i] == px[i0] && py[i] < py[i0])) i0 = i;
    int hull[8], m = 0, ih = i0;
    for (;;) {
        hull[m] = ih;
        int ie = 0;
        for (int j = 1; j < n; ++j) {
            if (ie == ih) { ie = j; continue; }
            float rx = px[ie] - px[hull[m]], ry = py[ie] - py[hull[m]];
            float vx = px[j] - px[hull[m]], vy = py[j] - py[hull[m]];
            float c = rx * vy - ry * vx;
            if (c < 0.0f) ie = j;
            if (c == 0.0f && (vx * vx + vy * vy) > (rx * rx + ry * ry)) ie = j;
        }
        ++m;
        ih = ie;
        if (ie == i0) break;
    }
    p->count = m;
    for (int i = 0; i < 8; ++i) { p->vx[i] = p->vy[i] = p->nx[i] = p->ny[i] = 0.0f; }
    for (int i = 0; i < m; ++i) { p->vx[i] = px[hull[i]]; p->vy[i] = py[hull[i]]; }
    for (int i = 0; i < m; ++i) {
        int i2 = i + 1 < m ? i + 1 : 0;
        float ex = p->vx[i2] - p->vx[i], ey = p->vy[i2] - p->vy[i];
        float nx = 1.0f * ey, ny = -1.0f * ex;  // b2Cross(edge, 1.0f)
        float len = sqrtf(nx * nx + ny * ny);
        float inv = 1.0f / len;
        p->nx[i] = nx * inv;
        p->ny[i] = ny * inv;
    }
}
struct HostMass {
    float mass, cx, cy, I;
};
inline void poly_mass(const HostPoly* p, float density, HostMass* md) {  // b2PolygonShape::ComputeMass
    float cx = 0.0f, cy = 0.0f, area = 0.0f, I = 0.0f;
    float sx = 0.0f, sy = 0.0f;
    for (int i = 0; i < p->count; ++i) { sx += p->vx[i]; sy += p->vy[i]; }
    float invn = 1.0f / p->count;
    sx *= invn; sy *= invn;
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < p->count; ++i) {
        int i2 = i + 1 < p->count ? i + 1 : 0;
        float e1x = p->vx[i] - sx, e1y = p->vy[i] - sy;
        float e2x = p->vx[i2] - sx, e2y = p->vy[i2] - sy;
        float D = e1x * e2y - e1y * e2x;
        float triangleArea = 0.5f * D;
        area += triangleArea;
        float f = triangleArea * k_inv3;
        cx += f * (e1x + e2x);
        cy += f * (e1y + e2y);
        float intx2 = e1x * e1x + e2x * e1x + e2x * e2x;
        float inty2 = e1y * e1y + e2y * e1y + e2y * e2y;
        I += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    md->mass = density * area;
    float inva = 1.0f / area;
    cx *= inva; cy *= inva;
    md->cx = cx + sx;
    md->cy = cy + sy;
    md->I = density * I;
    md->I += md->mass * ((md->cx * md->cx + md->cy * md->cy) - (cx * cx + cy * cy));
}
// b2Body::ResetMassData over fixtures given newest-first
struct HostBodyMass {
    float mass, invMass, I, invI, lcx, lcy;
};
inline void body_mass(const HostPoly* const* polys, const float* dens, int nfix, HostBodyMass* out) {
    float mass = 0.0f, I = 0.0f, lcx = 0.0f, lcy = 0.0f;
    for (int i = 0; i < nfix; ++i) {
        if (dens[i] == 0.0f) continue;
        HostMass md;
        poly_mass(polys[i], dens[i], &md);
        mass += md.mass;
        lcx += md.mass * md.cx;
        lcy += md.mass * md.cy;
        I += md.I;
    }
    float invMass, invI;
    if (mass > 0.0f) {
        invMass = 1.0f / mass;
        lcx *= invMass;
        lcy *= invMass;
    } else {
        mass = 1.0f;
        invMass = 1.0f;
    }
    if (I > 0.0f) {
        I -= mass * (lcx * lcx + lcy * lcy);
        invI = 1.0f / I;
    } else {
        I = 0.0f;
        invI = 0.0f;
    }
    out->mass = mass; out->invMass = invMass; out->I = I; out->invI = invI; out->lcx = lcx; out->lcy = lcy;
}

inline void put_shape(float* ctab, int s, const HostPoly* p) {
    float* t = ctab + CT_SHAPES + kShapeWords * s;
    t[0] = (float)p->count;
    for (int i = 0; i < 8; ++i) { t[1 + i] = p->vx[i]; t[9 + i] = p->vy[i]; t[17 + i] = p->nx[i]; t[25 + i] = p->ny[i]; }
    // bounding data for the exact culling tests (conservative: padded outwards)
    float lx = p->vx[0], hx = p->vx[0], ly = p->vy[0], hy = p->vy[0];
    for (int i = 1; i < p->count; ++i) {
        lx = p->vx[i] < lx ? p->vx[i] : lx; hx = p->vx[i] > hx ? p->vx[i] : hx;
        ly = p->vy[i] < ly ? p->vy[i] : ly; hy = p->vy[i] > hy ? p->vy[i] : hy;
    }
    float* x = ctab + CT_SHAPEX + 6 * s;
    const float pad = 1e-5f;
    x[0] = 0.5f * (lx + hx); x[1] = 0.5f * (ly + hy);
    x[2] = 0.5f * (hx - lx) + pad; x[3] = 0.5f * (hy - ly) + pad;
    float r2 = 0.0f;
    for (int i = 0; i < p->count; ++i) {
        float dx = p->vx[i] - x[0], dy = p->vy[i] - x[1];
        r2 = dx * dx + dy * dy > r2 ? dx * dx + dy * dy : r2;
    }
    x[4] = sqrtf(r2) + pad;
    x[5] = p->count == 4 ? 1.0f : 0.0f;  // every 4-gon here is an axis-aligned box in its local frame
}

// Fills the variant part of SimConst and the CT_WORDS-float constant table.
inline int build_variant(int variant, int n_agents, SimConst* K, float* ctab, mrp_layout* L) {
    int rc = mrp_layout_for(variant, n_agents, L);
    if (rc) return rc;
    // contact capacity is 32 slots per env (one 32-bit mask per flag set): enough for the registered
    // variants (measured max 21 over 1.2M env-steps) but not for v2 with > 2 three-fixture robots,
    // whose fat-AABB pair count averages 20 (n=3) .. 60 (n=5).  See DESIGN.md "Limits".
    if (L->max_contacts > kMaxC || L->n_dyn_fixtures > kMaxDynFix) return -3;  // needs the wide-capacity compilation (mrp_b200.cu)
    memset(ctab, 0, sizeof(float) * CT_WORDS);
    const bool square = variant == MRP_VARIANT_SQUARE_V2;   // three blocks, Heavy-v2 dynamics (extension, include/mrp_state.h)
    const bool v2 = variant >= 2, heavy = square || (variant & 1) != 0;
    const int n = L->n_agents;
    K->nblk = square ? MRP_SQUARE_BLOCKS : 1; K->nbf = square ? 5 : 2;
    K->variant = variant; K->v2 = v2 ? 1 : 0; K->n = n; K->nb = n + K->nblk;
    K->nfix = L->n_fixtures; K->ndynfix = L->n_dyn_fixtures; K->per_agent = v2 ? 3 : 1;
    K->maxc = L->max_contacts; K->obs_dim = L->obs_dim; K->act_dim = L->act_dim; K->max_steps = L->max_episode_steps;
    K->w_body = W_DIST + 2 * (n + 1);
    K->w_aabb = K->w_body + kBodyWords * K->nb;
    K->w_con = K->w_aabb + 4 * K->ndynfix;
    K->w_total = K->w_con + MRP_CONTACT_WORDS * K->maxc;
    K->smem_words = kDynFields * K->nb + 24 + 4 * K->ndynfix;
    K->h = (float)(1.0 / 50);                       // world.Step(1.0/FPS, ...) mrp00:428
    K->lin_k = 1.0f / (1.0f + K->h * 5.0f);         // DAMP / LINEAR_DAMP = 5.0
    K->ang_k = 1.0f / (1.0f + K->h * 5.0f);
    double SCALE, VW, VH;
    HostPoly stem, bar, oct, wheel1, wheel2, wallLR, wallBT;
    float blk_density, blk_friction, ag_density, ag_friction;
    if (!v2) {
        SCALE = 30.0; VW = 640; VH = 480;             // mrp00:40-42
        double S = 2.0, scaled = heavy ? S / 2 : S;   // mrp00:303-308
        blk_density = (float)(heavy ? 5.0 * 2 : 5.0);
        blk_friction = 0.999f;
        poly_box(&stem, (float)(1 / scaled), (float)(1 / scaled), 0.0f, (float)(-1 / scaled));
        poly_box(&bar, (float)(3 / scaled), (float)(1 / scaled), 0.0f, (float)(1 / scaled));
        const double P[8][2] = {{-0.5 / S, -1.5 / S}, {0.5 / S, -1.5 / S}, {1.5 / S, -0.5 / S}, {1.5 / S, 0.5 / S},
                                {0.5 / S, 1.5 / S},   {-0.5 / S, 1.5 / S}, {-1.5 / S, 0.5 / S}, {-1.5 / S, -0.5 / S}};
        float px[8], py[8];
        for (int k = 0; k < 8; ++k) { px[k] = (float)P[k][0]; py[k] = (float)P[k][1]; }
        poly_hull(&oct, px, py, 8);
        ag_density = 0.0f;   // fixtureDef default (mrp00:370-371)
        ag_friction = 0.2f;
        K->goal_x0 = 320.0 + 0.0 * SCALE;             // set_final_loc mrp00:115-128
        K->goal_y0 = 240.0 + 0.75 * SCALE;
        K->rp.agentDelta = 10; K->rp.agentDistance = 0.1; K->rp.blockDelta = 50; K->rp.blockDistance = 0.025;
    } else {
        SCALE = 140.0 * 4; VW = 1440; VH = 810;       // mrp02:40-43
        blk_density = (float)(heavy ? 20.0 : 1.56);
        blk_friction = 0.01f;
        poly_box(&stem, 0.1f, 0.1f, 0.0f, -0.1f);
        poly_box(&bar, 0.3f, 0.1f, 0.0f, 0.1f);
        const double P[8][2] = {{-0.039, -0.095}, {0.039, -0.095}, {0.095, -0.039}, {0.095, 0.039},
                                {0.039, 0.095},   {-0.039, 0.095}, {-0.095, 0.039}, {-0.095, -0.039}};
        float px[8], py[8];
        for (int k = 0; k < 8; ++k) { px[k] = (float)P[k][0]; py[k] = (float)P[k][1]; }
        poly_hull(&oct, px, py, 8);
        poly_box(&wheel1, 0.005f, 0.05f, 0.06f, 0.0f);
        poly_box(&wheel2, 0.005f, 0.05f, -0.06f, 0.0f);
        ag_density = 17.3f;
        ag_friction = 0.01f;
        K->goal_x0 = 0; K->goal_y0 = 0;
        K->rp.agentDelta = 10; K->rp.agentDistance = 0.25; K->rp.blockDelta = 25; K->rp.blockDistance = 0.1;
    }
    K->rp.puzzleComp = 10000; K->rp.outOfBounds = 1000; K->rp.blkOutOfBounds = 100;
    K->rp.scaled_epsilon = 0.1; K->rp.decay_pow = 1.0;
    K->SCALE = SCALE; K->W = VW / SCALE; K->H = VH / SCALE; K->ratio = SCALE / VW; K->SPEED = 10 / SCALE * 4;
    const double W = K->W, H = K->H;
    if (!v2) { poly_box_plain(&wallLR, 1.0f, (float)H); poly_box_plain(&wallBT, (float)W, 1.0f); }
    else     { poly_box_plain(&wallLR, 0.1f, (float)H); poly_box_plain(&wallBT, (float)W, 0.1f); }
    // mass: fixture list is newest first => bar, stem
    {
        const HostPoly* ps[2] = {&bar, &stem};
        float ds[2] = {blk_density, blk_density};
        HostBodyMass bm;
        body_mass(ps, ds, 2, &bm);
        K->blk_mass = bm.mass; K->blk_invMass = bm.invMass; K->blk_invI = bm.invI; K->blk_lcx = bm.lcx; K->blk_lcy = bm.lcy;
    }
    {
        HostBodyMass am;
        if (v2) {
            const HostPoly* ps[3] = {&wheel2, &wheel1, &oct};
            float ds[3] = {0.0f, 0.0f, ag_density};
            // ResetMassData runs when the octagon is attached; the density-0 wheels never re-run it
            body_mass(ps + 2, ds + 2, 1, &am);
        } else {
            am.mass = 1.0f; am.invMass = 1.0f; am.I = 0.0f; am.invI = 0.0f; am.lcx = 0.0f; am.lcy = 0.0f;  // b2Body ctor
        }
        K->ag_mass = am.mass; K->ag_invMass = am.invMass; K->ag_invI = am.invI; K->ag_lcx = am.lcx; K->ag_lcy = am.lcy;
        K->ag_inertia = am.I + am.mass * (am.lcx * am.lcx + am.lcy * am.lcy);  // b2Body::GetInertia
    }
    for (int i = 0; i < 4; ++i) {
        K->blkv[i][0] = bar.vx[i]; K->blkv[i][1] = bar.vy[i];
        K->blkv[4 + i][0] = stem.vx[i]; K->blkv[4 + i][1] = stem.vy[i];
    }
    K->blkv_off[0] = 0; K->blkv_off[1] = 8; K->blkv_off[2] = 8; K->blkv_off[3] = 8;
    K->blkx_invMass[0] = K->blk_invMass; K->blkx_invI[0] = K->blk_invI; K->blkx_lcx[0] = K->blk_lcx; K->blkx_lcy[0] = K->blk_lcy;
    // shapes: 0 stem, 1 bar, 2 octagon, 3 wheel1, 4 wheel2, 5 wall L/R, 6 wall B/T; square variant: 7 L small, 8 L tall, 9 I
    put_shape(ctab, 0, &stem); put_shape(ctab, 1, &bar); put_shape(ctab, 2, &oct);
    if (v2) { put_shape(ctab, 3, &wheel1); put_shape(ctab, 4, &wheel2); }
    put_shape(ctab, 5, &wallLR); put_shape(ctab, 6, &wallBT);
    if (square) {
        // L and I blocks (mrp00:334-351, blocks.py:92-109) at the v2 T-block's unit u = 0.1; creation order small box, tall box
        const float u = 0.1f;
        HostPoly lsmall, ltall, ibox;
        poly_box(&lsmall, 1 * u, 1 * u, 1 * u, 0.5f * u);
        poly_box(&ltall, 1 * u, 2 * u, -1 * u, -0.5f * u);
        poly_box_plain(&ibox, 1 * u, 2 * u);
        put_shape(ctab, 7, &lsmall); put_shape(ctab, 8, &ltall); put_shape(ctab, 9, &ibox);
        {   // fixture lists newest first
            const HostPoly* ps[2] = {&ltall, &lsmall};
            float ds[2] = {blk_density, blk_density};
            HostBodyMass bm;
            body_mass(ps, ds, 2, &bm);
            K->blkx_invMass[1] = bm.invMass; K->blkx_invI[1] = bm.invI; K->blkx_lcx[1] = bm.lcx; K->blkx_lcy[1] = bm.lcy;
            const HostPoly* pi[1] = {&ibox};
            body_mass(pi, ds, 1, &bm);
            K->blkx_invMass[2] = bm.invMass; K->blkx_invI[2] = bm.invI; K->blkx_lcx[2] = bm.lcx; K->blkx_lcy[2] = bm.lcy;
        }
        // observation vertices per block: fixtures newest first, vertices not listed before (mrp00:356-361): L has 7
        int nv = 8;
        const HostPoly* lists[2][2] = {{&ltall, &lsmall}, {&ibox, nullptr}};
        for (int k = 0; k < 2; ++k) {
            const int first = nv;
            for (int fi = 0; fi < 2 && lists[k][fi]; ++fi)
                for (int i = 0; i < 4; ++i) {
                    const float x = lists[k][fi]->vx[i], y = lists[k][fi]->vy[i];
                    bool seen = false;
                    for (int q = first; q < nv; ++q) if (K->blkv[q][0] == x && K->blkv[q][1] == y) seen = true;
                    if (!seen) { K->blkv[nv][0] = x; K->blkv[nv][1] = y; ++nv; }
                }
            K->blkv_off[2 + k] = nv;
        }
        // target poses of mrp00:83-88 (given there for a block unit of 0.5 m) at u = 0.1: the COMs of T, L (turned by pi/2) and I
        // that tile the square [-3u, 3u]^2 around the goal centre
        const double rel[3][3] = {{0.0, 0.75 / 0.5, 0.0}, {-2. / 3. / 0.5, -2. / 3. / 0.5, 0.5 * 3.14159265358979323846}, {1.0 / 0.5, -0.5 / 0.5, 0.0}};
        for (int k = 0; k < 3; ++k) { K->sq_target[k][0] = rel[k][0] * 0.1; K->sq_target[k][1] = rel[k][1] * 0.1; K->sq_target[k][2] = rel[k][2]; }
    }
    int f = 0;
    auto fix = [&](int body, int shape, float fr) {
        ctab[CT_FIXBODY + f] = (float)body; ctab[CT_FIXSHAPE + f] = (float)shape; ctab[CT_FIXFRIC + f] = fr; ++f;
    };
    fix(0, 0, blk_friction);
    fix(0, 1, blk_friction);
    if (square) { fix(1, 7, blk_friction); fix(1, 8, blk_friction); fix(2, 9, blk_friction); }
    for (int i = 0; i < n; ++i) {
        fix(K->nblk + i, 2, ag_friction);
        if (v2) { fix(K->nblk + i, 3, ag_friction); fix(K->nblk + i, 4, ag_friction); }
    }
    const double borders[4][2] = {{0, 0.5}, {1, 0.5}, {0.5, 0}, {0.5, 1}};
    for (int k = 0; k < 4; ++k) {
        float px = (float)(W * borders[k][0]), py = (float)(H * borders[k][1]);
        ctab[CT_WALLPOS + 2 * k] = px;
        ctab[CT_WALLPOS + 2 * k + 1] = py;
        const HostPoly* wp = k < 2 ? &wallLR : &wallBT;
        // ComputeAABB at xf = (p, angle 0) then +- aabbExtension (CreateProxy)
        float lx = 0, ly = 0, hx = 0, hy = 0;
        for (int i = 0; i < 4; ++i) {
            float x = (1.0f * wp->vx[i] - 0.0f * wp->vy[i]) + px, y = (0.0f * wp->vx[i] + 1.0f * wp->vy[i]) + py;
            if (i == 0) { lx = hx = x; ly = hy = y; }
            else { lx = x < lx ? x : lx; ly = y < ly ? y : ly; hx = x > hx ? x : hx; hy = y > hy ? y : hy; }
        }
        float* wb = ctab + CT_WALLBOX + 4 * k;
        wb[0] = lx; wb[1] = ly; wb[2] = hx; wb[3] = hy;
        float* wf = ctab + CT_WALLFAT + 4 * k;
        wf[0] = (lx - kPolygonRadius) - kAabbExtension; wf[1] = (ly - kPolygonRadius) - kAabbExtension;
        wf[2] = (hx + kPolygonRadius) + kAabbExtension; wf[3] = (hy + kPolygonRadius) + kAabbExtension;
        fix(K->nb + k, k < 2 ? 5 : 6, 0.2f);
    }
    for (int b = 0; b < K->nb; ++b) {   // walls (bodies nb .. nb + 3) keep the zeros of the memset
        float* bp = ctab + CT_BODYP + 4 * b;
        bp[0] = b == 0 ? K->blk_invMass : (b < K->nblk ? K->blkx_invMass[b] : K->ag_invMass);
        bp[1] = b == 0 ? K->blk_invI : (b < K->nblk ? K->blkx_invI[b] : K->ag_invI);
        bp[2] = b == 0 ? K->blk_lcx : (b < K->nblk ? K->blkx_lcx[b] : K->ag_lcx);
        bp[3] = b == 0 ? K->blk_lcy : (b < K->nblk ? K->blkx_lcy[b] : K->ag_lcy);
    }
    return 0;
}

}  // namespace mrp
