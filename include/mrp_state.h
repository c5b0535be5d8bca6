/*
 * mrp_state.h — canonical per-env state exchange record.
 *
 * Shared by the product C-ABI (include/mrp_b200.h: mrp_get_state / mrp_set_state)
 * and by the test oracle (oracle/oracle_capi.cpp: orc_get_state / orc_set_state) so
 * that the parity harness can load *identical states* into both, step once and
 * compare (BASELINE.json north_star: "Correctness is checked ... from identical
 * states").  It replaces nothing in the reference: pybox2d owns this state inside
 * SWIG C++ objects (b2Body / b2Contact / b2BroadPhase) and gym_puzzles never
 * serialises it (SURVEY.md §5 "Checkpoint / resume: none").
 *
 * The record is a flat array of 32-bit words (int32 / float32 / halves of float64),
 * one record per env, records contiguous (AoS, host or device memory).
 *
 * Layout for a variant with n robots, F_dyn fixtures on dynamic bodies and a
 * contact capacity MAXC (all three reported by mrp_layout; MAXC <= 32, or <= 192 for v2 with num_agents > 2):
 *
 *   word 0            i32  elapsed_steps   (TimeLimit counter, gym_puzzles/__init__.py:3-29)
 *   word 1            i32  episode         (Philox spawn counter)
 *   word 2            i32  blks_in_place   (mrp00:173-174,502-506 — persists across resets)
 *   word 3            i32  n_contacts
 *   word 4..11        i32  goal_contact[i] (agent i, 0/1; mrp00:92-111; slots >= n are 0)
 *   word 12..         f32  bodies[(n+1)][6] = {c.x, c.y, a, v.x, v.y, w}; body 0 = T-block,
 *                          body 1+i = agent i   (b2Body m_sweep.c / m_sweep.a / velocities)
 *                          [square variant: bodies 0..2 = T, L, I block, body 3+i = agent i; word 2 is then also the index
 *                           of the current goal block; fixtures 0 T stem, 1 T bar, 2 L small, 3 L tall, 4 I, then the robots]
 *   then              f64  agent_dist[n], block_dist          (mrp00:277-291 prev distances)
 *   then              f64  goal_x, goal_y                     (mrp02:303-311; v0: 320, 262.5)
 *   then              f64  ep_return ; i32 ep_len ; i32 pad
 *   then              f32  fat_aabb[F_dyn][4] = {lo.x, lo.y, hi.x, hi.y} (b2DynamicTree fat AABBs)
 *   then              contacts[MAXC][14], world contact-list order, head (newest) first:
 *        w0  u32  fixtureA | fixtureB<<8 | touching<<16 | type<<17 (0 faceA, 1 faceB) | pointCount<<18
 *        w1  u32  idkey(point0) | idkey(point1)<<16 ; idkey = indexA | indexB<<4 | typeA<<8 | typeB<<9
 *        w2,w3    manifold.localNormal     w4,w5  manifold.localPoint
 *        w6..w9   point0 {localPoint.x, localPoint.y, normalImpulse, tangentImpulse}
 *        w10..w13 point1 {...}
 *
 * Fixture indices are proxy creation order in a fresh b2World:
 *   0 stem, 1 bar, then each agent's fixtures (v0: octagon; v2: octagon, wheel1, wheel2),
 *   then wall_left, wall_right, wall_bottom, wall_top.
 */
#ifndef MRP_STATE_H
#define MRP_STATE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRP_MAX_AGENTS 8
#define MRP_CONTACT_WORDS 14

/* variant ids (gym ids of gym_puzzles/__init__.py:3-29) */
enum {
    MRP_VARIANT_V0 = 0,       /* MultiRobotPuzzle-v0       mrp00:142-150 */
    MRP_VARIANT_HEAVY_V0 = 1, /* MultiRobotPuzzleHeavy-v0  mrp00:606-610 */
    MRP_VARIANT_V2 = 2,       /* MultiRobotPuzzle-v2       mrp02:126-137 */
    MRP_VARIANT_HEAVY_V2 = 3, /* MultiRobotPuzzleHeavy-v2  mrp02:711-712 */
    /* BASELINE.json configs[4]: three blocks (T, L, I) that form a square, Heavy-v2 dynamics.  NOT a reference env: the
     * reference carries only its ingredients (L / I fixtures mrp00:334-351, blocks.py:92-109; target poses as comments
     * mrp00:83-88; block_queue / _set_next_goal_block mrp00:293-297).  Semantics are defined by oracle/mrp_env.hpp
     * ("square" section) and DESIGN.md §12; parity for it is oracle-defined. */
    MRP_VARIANT_SQUARE_V2 = 4 /* MultiRobotPuzzleSquare-v2 (extension) */
};
#define MRP_SQUARE_BLOCKS 3

typedef struct mrp_layout {
    int32_t n_agents;
    int32_t n_dyn_bodies;   /* n_agents + number of blocks (1; 3 for the square variant): blocks first, then the robots */
    int32_t n_fixtures;     /* incl. the 4 walls */
    int32_t n_dyn_fixtures; /* fixtures on dynamic bodies */
    int32_t max_contacts;   /* MAXC */
    int32_t obs_dim;
    int32_t act_dim;
    int32_t max_episode_steps;
    int32_t off_goal_contact; /* word offsets into the record */
    int32_t off_bodies;
    int32_t off_dists;
    int32_t off_goal;
    int32_t off_episode_acc;
    int32_t off_aabb;
    int32_t off_contacts;
    int32_t state_words;
} mrp_layout;

/* Pure function of (variant, n_agents): fills *out; returns 0 or a negative error.
 * n_agents <= 0 selects the registered default (2; 5 for Heavy-v0). */
static inline int mrp_layout_for(int variant, int n_agents, mrp_layout* L) {
    if (variant < 0 || variant > 4) return -1;
    int v2 = variant >= 2, square = variant == MRP_VARIANT_SQUARE_V2;
    if (n_agents <= 0) n_agents = (variant == MRP_VARIANT_HEAVY_V0) ? 5 : 2;
    if (n_agents > MRP_MAX_AGENTS || (square && n_agents > 7)) return -2;
    int per_agent = v2 ? 3 : 1;
    int nblk = square ? MRP_SQUARE_BLOCKS : 1, blk_fix = square ? 5 : 2;   /* T 2 + L 2 + I 1 */
    L->n_agents = n_agents;
    L->n_dyn_bodies = n_agents + nblk;
    L->n_dyn_fixtures = blk_fix + per_agent * n_agents;
    L->n_fixtures = L->n_dyn_fixtures + 4;
    /* potential contacts: fixture pairs on different bodies with >=1 dynamic body */
    int fa = per_agent * n_agents;
    int pot = blk_fix * fa + per_agent * per_agent * n_agents * (n_agents - 1) / 2 + 4 * L->n_dyn_fixtures + (square ? 8 : 0);
    /* capacity: 32 covers every registered variant (measured maximum 21 on long rollouts; overflow is counted);
     * MultiRobotPuzzle2(num_agents > 2) puts three-fixture robots side by side and reaches > 100 live fat-AABB pairs
     * (188 possible with 5 robots), so it gets the wide capacity */
    /* the square variant packs three multi-fixture blocks side by side: 32 slots overflow about once per 6e4 env-steps */
    int cap = ((v2 && n_agents > 2) || square) ? 192 : 32;
    L->max_contacts = pot < cap ? pot : cap;
    /* square: per robot 9, per block 4 + 2 * vertices (T 8, L 7 after the reference's de-duplication, I 4), epsilon,
     * goal-block index, blocks in place, robots in contact with the goal block */
    L->obs_dim = square ? 9 * n_agents + (4 + 16) + (4 + 14) + (4 + 8) + 4 : (v2 ? 9 * n_agents + 21 : 4 * n_agents + 20);
    L->act_dim = v2 ? 2 * n_agents : 3 * n_agents;
    L->max_episode_steps = (variant == MRP_VARIANT_HEAVY_V0) ? 3000 : 2000;
    int o = 4;
    L->off_goal_contact = o; o += MRP_MAX_AGENTS;
    L->off_bodies = o;       o += 6 * L->n_dyn_bodies;
    L->off_dists = o;        o += 2 * (n_agents + 1);
    L->off_goal = o;         o += 4;
    L->off_episode_acc = o;  o += 4;
    L->off_aabb = o;         o += 4 * L->n_dyn_fixtures;
    L->off_contacts = o;     o += MRP_CONTACT_WORDS * L->max_contacts;
    L->state_words = o;
    return 0;
}

#ifdef __cplusplus
}
#endif
#endif /* MRP_STATE_H */
