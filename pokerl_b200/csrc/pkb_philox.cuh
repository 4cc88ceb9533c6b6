// pkb_philox.cuh -- counter-based dealing and action sampling (Philox4x32-10).
//
// Replaces random.shuffle(self.deck) (reference game.py:424, CPython's global MT19937) with a
// pure function of (seed, env_id, episode, hand): any env's deck can be regenerated anywhere --
// at showdown inside the step kernel, for observations, or on the host to replay the hand into
// the reference Game.  Decks are therefore never stored in HBM.
//
// Deal spec: draw i of a hand (i = 0 .. 4+2P) takes 32-bit word (i & 3) of the Philox block with
// counter {env_id, hand, episode, i >> 2} and key {seed_lo, seed_hi}; j = mulhi32(word, 52 - i)
// selects the j-th smallest card id not drawn yet.  Draws 0..4 are the board (game.py:278),
// draws 5+2p, 6+2p the hole cards of seat p (game.py:385-389).
//
// Action spec: counter {env_id & ~96, step_idx, 0, 0x80000000}, word (env_id >> 5) & 3 -- the four envs whose
// ids differ only in bits 5 and 6 share one Philox block and take one word each, so a lane of the step kernel,
// which handles exactly such envs (slots lane, lane+32, lane+64, lane+96 of a 128-env tile), computes one
// block per tile instead of four.  t = mulhi32(word, W) with W = total weight of the valid moves; the move is the first valid one whose cumulative
// weight exceeds t (uniform weights = RandomAgent, agents/random.py:16-17).
#pragma once
#include "pkb_common.cuh"

namespace pkb {

struct U4 {
    uint32_t x, y, z, w;
};

// inlined form: several independent blocks of one thread interleave (the deal of a showdown)
PKB_HD U4 philox4x32_10_inl(U4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
PKB_HD_CALL U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) { return philox4x32_10_inl(c, k0, k1); }

// position of the j-th (0-based) set bit of a 52-bit mask; j < popcount(f)
PKB_HD uint32_t select_bit_inl(uint64_t f, uint32_t j)
{
    uint32_t w = (uint32_t)f, pos = 0;
    uint32_t c = (uint32_t)popc32(w);
    bool up = j >= c;
    j -= up ? c : 0u; w = up ? (uint32_t)(f >> 32) : w; pos = up ? 32u : 0u;
    c = (uint32_t)popc32(w & 0xffffu); up = j >= c;
    j -= up ? c : 0u; w = up ? w >> 16 : w; pos += up ? 16u : 0u;
    c = (uint32_t)popc32(w & 0xffu); up = j >= c;
    j -= up ? c : 0u; w = up ? w >> 8 : w; pos += up ? 8u : 0u;
    c = (uint32_t)popc32(w & 0xfu); up = j >= c;
    j -= up ? c : 0u; w = up ? w >> 4 : w; pos += up ? 4u : 0u;
    c = (uint32_t)popc32(w & 0x3u); up = j >= c;
    j -= up ? c : 0u; w = up ? w >> 2 : w; pos += up ? 2u : 0u;
    return pos + (j >= (w & 1u) ? 1u : 0u);
}
PKB_HD_CALL uint32_t select_bit(uint64_t f, uint32_t j)
{
    uint32_t w = (uint32_t)f, pos = 0;
    uint32_t c = (uint32_t)popc32(w);
    if (j >= c) { j -= c; w = (uint32_t)(f >> 32); pos = 32; }
    c = (uint32_t)popc32(w & 0xffffu);
    if (j >= c) { j -= c; w >>= 16; pos += 16; }
    c = (uint32_t)popc32(w & 0xffu);
    if (j >= c) { j -= c; w >>= 8; pos += 8; }
    c = (uint32_t)popc32(w & 0xfu);
    if (j >= c) { j -= c; w >>= 4; pos += 4; }
    c = (uint32_t)popc32(w & 0x3u);
    if (j >= c) { j -= c; w >>= 2; pos += 2; }
    c = w & 1u;
    if (j >= c) pos += 1;
    return pos;
}

// Sequential dealer: call next() 5+2P times; ids are card ids (0..51).
struct Dealer {
    uint32_t k0, k1, env, hand, episode;
    uint64_t free_mask;
    U4 blk;
    int i;
    PKB_HDM Dealer(uint64_t seed, uint32_t env_id, uint32_t episode_, uint32_t hand_)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), env(env_id), hand(hand_), episode(episode_),
          free_mask((1ull << 52) - 1), blk{0, 0, 0, 0}, i(0)
    {
    }
    PKB_HDM uint32_t next()
    {
        const int lane = i & 3;
        // inline and branch-free: out-of-line calls from code whose lanes have drifted apart run lane group by lane group
        if (lane == 0) blk = philox4x32_10_inl(U4{env, hand, episode, (uint32_t)(i >> 2)}, k0, k1);
        const uint32_t word = lane == 0 ? blk.x : lane == 1 ? blk.y : lane == 2 ? blk.z : blk.w;
        const uint32_t id = select_bit_inl(free_mask, mulhi32(word, (uint32_t)(52 - i)));
        free_mask &= ~(1ull << id);
        ++i;
        return id;
    }
};

// The same deal for small tables, all draws at once: the N = 5+2P card ids of (seed, env, episode, hand).
// The Philox blocks are independent and interleave; the draw chain keeps the ids drawn so far SORTED as
// t[j] = (j-th smallest id) - j (non-decreasing), so that "the r-th smallest id not drawn yet" is
// r + #{j : t[j] <= r} -- independent compares instead of a six-level popcount search per draw.
template <int N>
PKB_HD void deal_ids(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t hand, uint32_t (&id)[N])
{
    constexpr int NB = (N + 3) / 4;
    uint32_t word[NB * 4];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const U4 blk = philox4x32_10_inl(U4{env_id, hand, episode, (uint32_t)b}, (uint32_t)seed, (uint32_t)(seed >> 32));
        word[4 * b] = blk.x; word[4 * b + 1] = blk.y; word[4 * b + 2] = blk.z; word[4 * b + 3] = blk.w;
    }
    uint32_t t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const uint32_t r = mulhi32(word[i], (uint32_t)(52 - i));
        uint32_t k = 0;
#pragma unroll
        for (int j = 0; j < i; ++j) k += t[j] <= r ? 1u : 0u;
        id[i] = r + k;
        // insert at position k: the entries from k on move up one place, and their value drops by one
#pragma unroll
        for (int j = i; j >= 0; --j) {
            const uint32_t below = j > 0 ? t[j > 0 ? j - 1 : 0] - 1u : 0u;
            t[j] = (uint32_t)j > k ? below : (uint32_t)j == k ? r : t[j];
        }
    }
}

// the 32-bit word that drives the action draw of (env, step): a pure function, independent of the state
PKB_HD U4 sample_block(uint64_t seed, uint32_t env_id, uint32_t step_idx)  // shared by env_id, env_id ^ 32, ^ 64, ^ 96
{
    return philox4x32_10(U4{env_id & ~96u, step_idx, 0u, 0x80000000u}, (uint32_t)seed, (uint32_t)(seed >> 32));
}
PKB_HD uint32_t sample_word_of(const U4& b, uint32_t env_id)
{
    const uint32_t k = env_id >> 5 & 3u;
    return k == 0 ? b.x : k == 1 ? b.y : k == 2 ? b.z : b.w;
}
PKB_HD uint32_t sample_word(uint64_t seed, uint32_t env_id, uint32_t step_idx)
{
    return sample_word_of(sample_block(seed, env_id, step_idx), env_id);
}

// weighted choice over the valid mask: the first valid move whose cumulative weight exceeds t
PKB_HD int pick_weighted(uint32_t word, uint32_t valid_mask, const uint32_t* weights)
{
    uint32_t total = 0;
#pragma unroll
    for (int a = 0; a < NUM_MOVES; ++a)
        if (valid_mask >> a & 1u) total += weights[a];
    const uint32_t t = mulhi32(word, total);
    uint32_t acc = 0;
    int pick = MV_FOLD;
    bool found = false;
#pragma unroll
    for (int a = 0; a < NUM_MOVES; ++a)
        if (valid_mask >> a & 1u) {
            acc += weights[a];
            if (!found && acc > t) { pick = a; found = true; }
        }
    return pick;
}

// uniform weights: the t-th valid move, t = mulhi32(word, number of valid moves) -- the position of the t-th set bit of
// the 7-bit mask, found by halving (low 4 bits, then 2, then 1)
PKB_HD int pick_uniform(uint32_t word, uint32_t valid_mask)
{
    valid_mask |= 1u;  // FOLD is always valid (game.py:367): an empty (never computed) mask folds instead of running off the end
    uint32_t t = mulhi32(word, (uint32_t)popc32(valid_mask)), w = valid_mask;
    int pick = 0;
    uint32_t c = (uint32_t)popc32(w & 0xfu);
    if (t >= c) { t -= c; w >>= 4; pick = 4; }
    c = (uint32_t)popc32(w & 0x3u);
    if (t >= c) { t -= c; w >>= 2; pick += 2; }
    if (t >= (w & 1u)) pick += 1;
    return pick;
}

// weights == nullptr -> uniform
PKB_HD int sample_action(uint64_t seed, uint32_t env_id, uint32_t step_idx, uint32_t valid_mask, const uint32_t* weights)
{
    const uint32_t word = sample_word(seed, env_id, step_idx);
    return weights ? pick_weighted(word, valid_mask, weights) : pick_uniform(word, valid_mask);
}

}  // namespace pkb
