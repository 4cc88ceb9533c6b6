"""Step 2 of the at-size parity chain (build container only: needs the live reference at /root/reference).

Replays the recordings of tests/golden/dump_gpu_at_size.py -- sampled envs of the FULL-SIZE runs of BASELINE.json
configs[2] / configs[3] on a B200 -- into the unmodified reference `pokerl.Game`: the env's decks are regenerated with
the host deal entry point of the product library (pkb_deal, no GPU needed) and injected through the DeckShim of
make_golden.py; every recorded action is played; the reference's (game_over, hand_over, round_over), its
get_valid_actions() and, at the end, every state field are compared bit for bit with what the GPU produced.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/replay_at_size.py gpurun_out     # -> tests/golden/at_size/*.npz + replay.log
"""
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (imports the live reference, DeckShim, snapshots)
from pokerl_b200 import _native as nat  # noqa: E402  (host-side pkb_deal only)


class ProductDeckShim(mg.DeckShim):
    """DeckShim dealing with the PRODUCT's host deal (pkb_deal) for a given global env id."""

    def __init__(self, seed, num_players, env_id):
        super().__init__(seed, num_players)
        self.env_id = env_id

    def shuffle(self, deck):
        self.in_step += 1
        if self.in_step > 40:
            raise mg.NonTerminating()
        first = nat.deal(self.seed, self.env_id, self.episode, self.game.hand, self.P)
        rest = [v for v in mg.ALL_VALUES if v not in set(first)]
        deck[:] = [mg.Card(v) for v in first + rest]


def replay(path, log):
    z = np.load(path)
    P, N, seed, T = [int(x) for x in z["config"]]
    start = z["start_credits"]
    start_arg = int(start[0]) if np.all(start == start[0]) else [float(x) for x in start]
    ids = z["env_ids"]
    bad = poisoned = steps = hands = games = 0
    t0 = time.time()
    for k, env in enumerate(ids):
        shim = ProductDeckShim(seed, P, int(env))
        mg.ref_game.random = shim
        mg.ref_game.np = mg.StableArgsortNumpy(np)
        g = mg.pokerl.Game(num_players=P, start_credits=start_arg, big_blind=2, small_blind=1)
        shim.game = g
        shim.episode += 1
        g.reset()
        ok = mg.valid_mask_of(g) == int(z["valid0"][k])
        dead = False
        for t in range(T):
            a, f_gpu, v_gpu = int(z["actions"][t, k]), int(z["flags"][t, k]), int(z["valid"][t, k])
            if f_gpu & nat.F_POISONED:
                dead = True  # the reference would spin or assert here: nothing to compare from this step on
                break
            shim.in_step = 0
            done, hand_over, round_over = g.step(a)
            steps += 1
            hands += int(bool(hand_over))
            flags = (1 if done else 0) | (2 if hand_over else 0) | (4 if round_over else 0)
            if done:  # the GPU batch runs with auto_reset: Game.reset() inside the step that ends the game
                games += 1
                shim.episode += 1
                g.reset()
            if flags != f_gpu or mg.valid_mask_of(g) != v_gpu:
                ok = False
                log(f"  MISMATCH {os.path.basename(path)} env {env} step {t}: reference flags {flags} valid {mg.valid_mask_of(g)}, GPU {f_gpu} {v_gpu}")
                break
        if dead:
            poisoned += 1
            continue
        if ok:  # final state, bit for bit
            s = mg.ref_snapshot(g)
            for key, name in (("credits", "credits"), ("bets", "bets"), ("pending", "pending"), ("payoffs", "payoffs")):
                ok &= np.array_equal(s[key], z["final_" + name][:, k])
            ok &= np.array_equal(s["states"], z["final_states"][:, k])
            ok &= int(s["minraise"]) == int(z["final_minraise"][k])
            for key in ("turn", "hand", "dealer", "sb", "bb", "active"):
                ok &= int(s[key]) == int(z["final_" + key][k])
            ok &= shim.episode == int(z["final_episode"][k])
            if not ok:
                log(f"  MISMATCH {os.path.basename(path)} env {env}: final state")
        bad += 0 if ok else 1
    log(f"{os.path.basename(path)}: P={P}, {N} envs on the GPU, {len(ids)} sampled envs x {T} steps replayed into the live reference: "
        f"{steps} reference steps, {hands} hand ends, {games} games, {poisoned} envs skipped (flagged non-terminating), "
        f"{bad} mismatching envs ({time.time() - t0:.0f} s)")
    return bad


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    out = os.path.join(HERE, "at_size")
    os.makedirs(out, exist_ok=True)
    lines = []

    def log(msg):
        print(msg, flush=True)
        lines.append(msg)

    log(f"replay_at_size.py: reference = sneppy/pokerl {getattr(mg.pokerl, '__version__', '0.0.6')} at /root/reference, numpy {np.__version__}, "
        "np.argsort pinned to the stable order (see make_golden.py)")
    total = 0
    for name in ("at_size_p2.npz", "at_size_p9.npz"):
        total += replay(os.path.join(src, name), log)
        shutil.copy(os.path.join(src, name), os.path.join(out, name))  # the validated recording becomes a fixture
    log("RESULT: " + ("bit-exact" if total == 0 else f"{total} mismatching envs"))
    open(os.path.join(out, "replay.log"), "w").write("\n".join(lines) + "\n")
    sys.exit(1 if total else 0)
