/*
 * pokerl_b200.h -- C ABI of the B200-native batched No-Limit Texas Hold'em hot path.
 *
 * Drop-in boundary for sneppy/pokerl's Game.step / dealing / judger path.  The
 * reference has no FFI of its own (it is pure Python + numpy); each entry point
 * below names the reference interface it replaces (file:line relative to the
 * reference tree).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every call returns 0 on success, < 0 on error; pkb_last_error() gives the text
 *   - no exceptions cross the ABI; no torch types; device pointers are plain void*
 *   - the handle owns the per-env device state; the caller owns action/output buffers
 *   - all launches are ordered on the cudaStream_t passed as `stream` (void*, may be NULL)
 *   - a handle is confined to one device and one host thread at a time; its calls may move between streams (a call
 *     on a new stream is ordered after everything the handle submitted to the previous one)
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *
 * Encodings (identical to the reference)
 *   card value  = (suit << 4) | rank0, rank0 0=Ace,1=Two..12=King      cards.py:4-26
 *   card id     = rank0 + 13*suit (bit index of 52-bit hand masks)      cards.py:22-26
 *   player state FOLDED 0, ACTIVE 1, CALLED 2, ALL_IN 3, BROKEN 4       enums.py:130-134
 *   move        FOLD 0, CHECK 1, CALL 2, RAISE_TEN 3, RAISE_QUARTER 4,
 *               RAISE_HALF 5, ALL_IN 6                                   enums.py:104-111
 *   hand rank   packed = (category << 20) | get_kickers_value(kickers)   judger.py:101-109
 *               category 1 = straight flush .. 9 = high card, 10 = none  enums.py:9-18
 */
#ifndef POKERL_B200_H
#define POKERL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PKB_ABI_VERSION 4
#define PKB_MAX_PLAYERS 10
#define PKB_NUM_MOVES 7

/* step flag bits written to flags[env] */
#define PKB_F_GAME_OVER 1u   /* Game.step()[0], game.py:565, 619 */
#define PKB_F_HAND_OVER 2u   /* Game.step()[1] */
#define PKB_F_ROUND_OVER 4u  /* Game.step()[2] */
#define PKB_F_SKIPPED 32u    /* the action was PKB_ACTION_SKIP: env untouched, not counted (masked lockstep advance) */
#define PKB_F_INVALID 64u    /* action not in valid_actions: env untouched (ValueError at game.py:649-651) */
#define PKB_F_POISONED 128u  /* reference would not terminate (game.py:607) or would assert (game.py:473) */

/* pkb_step action value that leaves an env alone: lets a caller advance only SOME envs of a batch, e.g. the
 * "step the opponents until the agent's seat is active" loops of PokerGameEnv (envs/game_env.py:24-27, 41-52) */
#define PKB_ACTION_SKIP (-2147483647 - 1)

#define PKB_NEXT_TURN_CAP 64 /* next_turn() calls allowed inside one step before PKB_F_POISONED */

typedef struct pkb_env pkb_env; /* opaque handle: N lockstep Game objects on one GPU */

/* Game(**config), game.py:242-264.  start_credits: one value per seat (game.py:247, 408). */
typedef struct pkb_config {
    uint32_t abi_version;     /* PKB_ABI_VERSION */
    int32_t device;           /* CUDA device ordinal */
    uint64_t num_envs;        /* N local environments */
    uint64_t env_id_base;     /* global id of local env 0 (multi-GPU sharding); global ids must be < 2^32 */
    int32_t num_players;      /* 2..PKB_MAX_PLAYERS                                  game.py:246 */
    int32_t dealer;           /* reset(dealer=...), default 0                        game.py:403 */
    double start_credits[PKB_MAX_PLAYERS]; /*                                        game.py:247 */
    double big_blind;         /*                                                     game.py:248 */
    double small_blind;       /*                                                     game.py:249 */
    uint64_t seed;            /* Philox key; replaces the global MT19937 of game.py:424 */
    int32_t auto_reset;       /* 1: reset() an env inside the step that ends its game */
    int32_t reserved;
} pkb_config;

/* Device pointers into the handle's state, for zero-copy wrapping (torch / cupy / raw CUDA).
 * credits / bets / pending_bets are [num_players][stride] (seat-major, env fastest), payoffs is [stride][payoff_pitch],
 * the rest are [stride]. */
typedef struct pkb_state_view {
    uint64_t num_envs;
    uint64_t stride;          /* elements between consecutive seats (num_envs rounded up to the tile) */
    int32_t num_players;
    int32_t payoff_pitch;     /* doubles per env in `payoffs`: num_players rounded up to a multiple of 4 */
    double *credits;          /* Game.credits                 game.py:260 */
    double *bets;             /* Game.bets                    game.py:261 */
    double *pending_bets;     /* Game.pending_bets            game.py:262 */
    double *payoffs;          /* Game.payoffs                 game.py:264 -- ENV-major, unlike the other planes:
                               * payoffs[env * payoff_pitch + seat]; the padding doubles of a record are 0.0.  (A step
                               * writes payoffs only where a hand ended; whole 32-byte records keep those scattered
                               * writes from pulling the rest of each sector out of HBM.) */
    double *minimum_raise_value; /* Game.minimum_raise_value  game.py:263 */
    uint64_t *seat_masks;     /* Game.player_states game.py:259 as seat bit masks, one word per env [stride]:
                               * ACTIVE | CALLED<<16 | ALL_IN<<32 | BROKEN<<48, bit p of a field = seat p (FOLDED: in
                               * none); pkb_player_states() gives one PlayerState value per seat */
    uint32_t *meta;           /* turn | dealer<<3 | small_blind_idx<<8 | big_blind_idx<<13 | active_player<<18 |
                               * valid<<23 (the active player's valid-move mask, cached by the step that produced it) */
    uint32_t *hand;           /* Game.hand                    game.py:255 */
    uint32_t *episode;        /* number of reset() calls seen by the env (Philox counter word) */
} pkb_state_view;

/* Aggregate counters kept on the device (pkb_stats). */
typedef struct pkb_stats_t {
    uint64_t steps;           /* valid env-steps executed */
    uint64_t hands;           /* end_hand() calls (game.py:453) */
    uint64_t showdowns;       /* end_hand() calls that ranked hands (game.py:483-528) */
    uint64_t games;           /* steps that returned game_over */
    uint64_t invalid;         /* steps rejected as invalid */
    uint64_t poisoned;        /* steps that hit the loop cap / the reference's assert */
    double payoff_sum[PKB_MAX_PLAYERS]; /* sum over hands of Game.payoffs per seat */
} pkb_stats_t;

const char *pkb_last_error(void);
int pkb_abi_version(void);
/* number of CUDA devices visible; < 0 if the runtime cannot initialise */
int pkb_device_count(void);

/* ---- Game: game.py ------------------------------------------------------------ */
int pkb_create(const pkb_config *cfg, pkb_env **out);                     /* Game.__init__  game.py:242 */
int pkb_destroy(pkb_env *h);
/* Game.reset, game.py:397-412.  env_mask: device u8[N], nonzero = reset that env; NULL = all. */
int pkb_reset(pkb_env *h, const uint8_t *env_mask, void *stream);
/* The `dealer` keyword of Game.reset(dealer=...), game.py:403: used by every later reset (and auto-reset). */
int pkb_set_reset_dealer(pkb_env *h, int32_t dealer);
/* Recomputes the valid-move mask cached in `meta` (bits 23-29) from the state planes.  pkb_step_random draws from the
 * cached mask; call this after writing credits / pending_bets / minimum_raise_value / meta through pkb_state().
 * (pkb_step and pkb_step_u8 always check the action against the planes as they are, game.py:648.) */
int pkb_refresh_valid(pkb_env *h, void *stream);
/* Consecutive step launches of one stream normally overlap: a launch starts while its predecessor drains, and tiles are
 * handed over one by one (programmatic dependent launch; results are identical).  enabled = 0 restores plain stream
 * order between step launches (measurement, debugging). */
int pkb_set_pipelining(pkb_env *h, int32_t enabled);
/* Game.step, game.py:621-700 on every env.  actions: device i32[N].  flags/valid: device u8[N] out
 * (valid = bit a set iff move a is valid for the NEW active player, game.py:339-383); either may be NULL. */
int pkb_step(pkb_env *h, const int32_t *actions, uint8_t *flags, uint8_t *valid, void *stream);
/* Same with one byte per action (a move index 0..6; PKB_ACTION_SKIP_U8 leaves the env alone; any other value is
 * rejected like an out-of-range int, game.py:649-651): a quarter of the host-to-device bytes of pkb_step for callers
 * whose actions come from the host every step. */
#define PKB_ACTION_SKIP_U8 255
int pkb_step_u8(pkb_env *h, const uint8_t *actions, uint8_t *flags, uint8_t *valid, void *stream);
/* The same step in its most compact host-facing form, for callers that ship actions up and results down every step:
 * actions_u4 = device u8[(N+1)/2], two move indices per byte (env 2i in the low nibble, env 2i+1 in the high one;
 * PKB_ACTION_SKIP_U4 leaves the env alone, 7..14 are rejected); packed_out = device u8[N], one byte per env:
 * bits 0-2 = PKB_F_GAME_OVER | PKB_F_HAND_OVER | PKB_F_ROUND_OVER, bits 3-7 = valid bits of moves 1..5 (CHECK, CALL,
 * RAISE_TEN, RAISE_QUARTER, RAISE_HALF) for the NEW active player -- FOLD and ALL_IN are always valid (game.py:367).
 * Rejected / poisoned steps do not fit the byte: they are counted (pkb_stats: invalid, poisoned) and leave bits 0-2
 * zero.  1.5 bytes per env-step cross PCIe instead of 3 (pkb_step_u8) or 6 (pkb_step). */
#define PKB_ACTION_SKIP_U4 15
int pkb_step_packed(pkb_env *h, const uint8_t *actions_u4, uint8_t *packed_out, void *stream);
/* Same, with the action drawn on device: weighted choice over the valid mask from
 * Philox(seed, env_id, step_idx) (RandomAgent, agents/random.py:16-17 when weights == NULL).
 * weights: host u32[7] or NULL.  actions_out: device i32[N] or NULL (the actions taken, for replay). */
int pkb_step_random(pkb_env *h, uint32_t step_idx, const uint32_t *weights, int32_t *actions_out, uint8_t *flags,
                    uint8_t *valid, void *stream);
/* ---- PokerGameEnv: envs/game_env.py ---------------------------------------------------------------------------
 * The reference's RL environment seats the learning agent in seat 0 and lets callables play the other seats.  With the
 * device-side random agent in those seats (weights as for pkb_step_random; NULL = RandomAgent, agents/random.py:16-17)
 * both of its entry points are ONE call each: the kernel keeps stepping an env -- one sampled opponent move per round,
 * inside the tile -- until the reference's loops would return.  The j-th opponent move of an env inside a call draws
 * with sampler step index round_base + j - 1 (j = 1, 2, ...; at most PKB_ENV_ROUND_CAP per call, then the env is given
 * up as poisoned), so callers advance round_base by PKB_ENV_ROUND_CAP from call to call.
 *
 * pkb_env_step = PokerGameEnv.step (game_env.py:31-54): actions = device i32[N], seat 0's move per env (PKB_ACTION_SKIP:
 * leave the env alone); reward = device f64[N] (game_env.py:39, 47; 0.0 for envs not stepped); done = device u8[N]:
 * bit 0 done, bit 1 hand_over, PKB_F_SKIPPED / PKB_F_INVALID / PKB_F_POISONED as in step flags (a poisoned env reports
 * done | hand_over so that the caller resets it); valid = device u8[N], the valid mask of the env's active player
 * afterwards (seat 0 unless the episode is over).  Envs are NOT reset on done (the caller does, as with the reference).
 * pkb_env_reset = PokerGameEnv.reset (game_env.py:20-29) for the envs selected by env_mask (device u8[N], NULL = all):
 * Game.reset, then the opponents act until it is seat 0's turn; a game that ends on the way is reset again (:27). */
#define PKB_ENV_ROUND_CAP 1024
int pkb_env_step(pkb_env *h, const int32_t *actions, uint32_t round_base, const uint32_t *weights, double *reward,
                 uint8_t *done, uint8_t *valid, void *stream);
int pkb_env_reset(pkb_env *h, const uint8_t *env_mask, uint32_t round_base, const uint32_t *weights, uint8_t *valid,
                  void *stream);
/* pkb_env_step on large handles runs in PHASES: the launch over the handle's planes plays three opponent rounds, the envs
 * whose opponents are still playing after that (an agent that has folded or gone broke; envs/game_env.py:41-52) are
 * packed into a scratch arena and stepped there by follow-up launches (6, 12, 24, then all remaining rounds).  Results do
 * not depend on it.  min_envs: handles with at least this many envs use phases (default 16384; 0: always; < 0: never).
 * The scratch arena (the size of the handle's state) is allocated by the first phased call. */
int pkb_set_env_phases(pkb_env *h, int64_t min_envs);
/* n_steps consecutive pkb_step_random steps (sampler step indices step_idx .. step_idx + n_steps - 1) in ONE launch: every
 * tile of envs is loaded once, stepped n_steps times in shared memory and stored once, so the state crosses HBM once per
 * n_steps Game.steps.  The state, the statistics and the payoffs afterwards are exactly those of n_steps single calls;
 * actions_out / flags / valid (each may be NULL) receive the LAST step's values.  For rollouts with device-side agents
 * whose intermediate states nobody reads. */
int pkb_step_random_n(pkb_env *h, uint32_t step_idx, uint32_t n_steps, const uint32_t *weights, int32_t *actions_out,
                      uint8_t *flags, uint8_t *valid, void *stream);
/* Sampler alone (K5): actions for the CURRENT state, without stepping. */
int pkb_sample_actions(pkb_env *h, uint32_t step_idx, const uint32_t *weights, int32_t *actions_out, void *stream);
/* Game.get_valid_actions(player), game.py:339-383.  player < 0: each env's active player. */
int pkb_valid_actions(pkb_env *h, int32_t player, uint8_t *valid, void *stream);
/* Game.game_over, game.py:316-320: device u8[N]. */
int pkb_game_over(pkb_env *h, uint8_t *out, void *stream);
int pkb_state(pkb_env *h, pkb_state_view *out);
/* Game.player_states, game.py:259, one PlayerState value per seat: out = device u8[num_players][stride] */
int pkb_player_states(pkb_env *h, uint8_t *out, void *stream);
/* Current deal of every env, device u8 [5+2P][stride]: rows 0..4 board (game.py:278),
 * rows 5+2p, 6+2p hole cards of seat p (game.py:385-389); card values. */
int pkb_cards(pkb_env *h, uint8_t *out, void *stream);
/* Observation features for q_learning.py:29-33: packed eval_hand(player_hand) of `player`
 * (< 0: active player) on 2/5/6/7 cards depending on the street.  out: device u32[N]. */
int pkb_hand_features(pkb_env *h, int32_t player, uint32_t *out, void *stream);
int pkb_stats(pkb_env *h, pkb_stats_t *out, void *stream);     /* synchronises the stream */
int pkb_stats_reset(pkb_env *h, void *stream);

/* ---- dealing: replaces random.shuffle(self.deck), game.py:424 ------------------ */
/* Host-side replay of one deal: out[5+2P] card values.  Pure function of its arguments. */
int pkb_deal(uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t hand, int32_t num_players, uint8_t *out);

/* ---- judger: judger.py --------------------------------------------------------- */
/* eval_hand on n hands of k cards each (0 <= k <= 7), judger.py:7-99.  cards: device u8[n][k]
 * card values (duplicates allowed, as in the reference's own tests).  out: device u32[n] packed. */
int pkb_eval_cards(const uint8_t *cards, int32_t k, uint64_t n, uint32_t *out, void *stream);
/* eval_hand on n hands given as 52-bit id masks (<= 7 bits set).  masks: device u64[n].  (Batches of >= 2^17 hands in
 * 16-byte aligned arrays take the streaming evaluator -- a perfect hash of the rank multiset held in shared memory, the
 * few hands it does not cover ranked in a second pass per warp; the result is the same word either way.) */
int pkb_eval_masks(const uint64_t *masks, uint64_t n, uint32_t *out, void *stream);
/* All C(52,k) hands (k = 1..7), enumerated on device in colex order of the id mask.
 * hist: host u64[11] by category; sums: host u64[2] = {sum packed, sum packed*(mask % 1000003)} mod 2^64;
 * out: device u32[count] or NULL (packed of hands first..first+count-1). */
int pkb_eval_exhaustive(int32_t k, uint64_t first, uint64_t count, uint64_t *hist, uint64_t *sums, uint32_t *out,
                        void *stream);
/* n random 7-card hands from Philox(seed, index): writes masks (device u64[n], may be NULL). */
int pkb_random_hands(uint64_t seed, uint64_t n, uint64_t *masks, void *stream);
/* compare_rankings, judger.py:111-158, on n tables of p packed rankings each.
 * packed: device u32[n][p]; onehot: device u8[n][p]. */
int pkb_compare_rankings(const uint32_t *packed, int32_t p, uint64_t n, uint8_t *onehot, void *stream);

#ifdef __cplusplus
}
#endif
#endif
