"""PokerGameEnv (pokerl/envs/game_env.py:6-54) on the GPU, scalar and batched, against the reference's
control flow restated over the oracle's scalar Game."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SKIP = -2 ** 31


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


class RefEnvOnOracle:
    """game_env.py:6-54 line by line, over oracle.OracleGame.  Opponents are callables valid_mask -> action."""

    def __init__(self, ora, agents, **cfg):
        self.o = ora.OracleGame(cfg["num_players"], cfg["start_credits"], cfg["big_blind"], cfg["small_blind"],
                                seed=cfg["seed"], env_id=cfg.get("env_id", 0))
        self.agents = [None, *agents]

    @property
    def active(self):
        return self.o.g.active_player

    def reset(self):  # game_env.py:20-29
        self.o.reset()
        while self.active != 0:
            f = self.o.step(self.agents[self.active](self.o.valid_mask()))
            if f & 1:
                self.o.reset()

    def step(self, action):  # game_env.py:31-54
        reward = .0
        f = self.o.step(action)
        assert not f & 64, "agent action rejected by the oracle"
        done, hand = bool(f & 1), bool(f & 2)
        if done or self.o.g.states[0] == 4:
            return self.o.g.payoffs[0], True, True
        while not hand and self.active != 0:
            f = self.o.step(self.agents[self.active](self.o.valid_mask()))
            done, hand = bool(f & 1), bool(f & 2)
        if hand:
            reward = self.o.g.payoffs[0]
        while not done and self.active != 0:
            f = self.o.step(self.agents[self.active](self.o.valid_mask()))
            done = bool(f & 1)
        return reward, done, hand


def bits(x):
    return np.float64(x).view(np.uint64)


def test_scalar_env_follows_the_reference_loops(torch_cuda, ora):
    from pokerl_b200.envs import PokerGameEnv
    cfg = dict(num_players=4, start_credits=60, big_blind=4, small_blind=2, seed=5)
    rng_gpu, rng_ref, rng_agent = np.random.default_rng(1), np.random.default_rng(1), np.random.default_rng(2)

    def gpu_opponent(state):
        return int(rng_gpu.choice(list(state.valid_action_indices)))

    def ref_opponent(mask):
        return int(rng_ref.choice([a for a in range(7) if mask >> a & 1]))

    env = PokerGameEnv(agents=[gpu_opponent] * 3, **cfg)
    ref = RefEnvOnOracle(ora, [ref_opponent] * 3, **cfg)
    state = env.reset()
    ref.reset()
    episodes = 0
    for t in range(400):
        idx = list(state.valid_action_indices)
        assert sum(1 << a for a in idx) == ref.o.valid_mask(), t
        a = int(rng_agent.choice(idx))
        state, reward, done, hand = env.step(a)
        r_ref, d_ref, h_ref = ref.step(a)
        assert (bits(reward), bool(done), bool(hand)) == (bits(r_ref), d_ref, h_ref), t
        snap = ref.o.snapshot()
        assert [bits(x) for x in env.game.credits] == [bits(x) for x in snap["credits"]], t
        assert env.game.active_player == snap["active"] and env.game.hand == snap["hand"], t
        if done:
            episodes += 1
            state = env.reset()
            ref.reset()
    assert episodes >= 3


def test_batched_env_equals_per_table_reference_flow(torch_cuda, ora):
    """N tables in lockstep with device-side RandomAgents; every table's trajectory is then replayed through the
    reference's scalar control flow (opponent decisions taken from the log of the launches)."""
    torch = torch_cuda
    from pokerl_b200.agents import RandomAgent
    from pokerl_b200.envs import PokerGameEnv
    N, P, T = 384, 3, 60
    cfg = dict(num_players=P, start_credits=40, big_blind=4, small_blind=2, seed=9)
    env = PokerGameEnv(agents=[RandomAgent() for _ in range(P - 1)], num_envs=N, fused=False, **cfg)  # the masked lockstep loops
    log = []  # (kind, actions [N]) per launch, in order
    inner = env.game.batch.step

    def logged_step(actions, *a, **k):
        log.append(actions.cpu().numpy().copy())
        return inner(actions, *a, **k)

    env.game.batch.step = logged_step
    gen = torch.Generator(device="cuda").manual_seed(3)
    state = env.reset()
    marks = [("reset", len(log), None)]
    for t in range(T):
        assert bool((env.game.batch.active_player == 0).all())
        a = torch.multinomial(state.valid_actions, 1, generator=gen)[:, 0].to(torch.int32)
        first = len(log)
        state, reward, done, hand = env.step(a)
        marks.append(("step", first, (a.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), hand.cpu().numpy())))
        if bool(done.any()):
            first = len(log)
            state = env.reset(env_mask=done)
            marks.append(("reset", first, done.cpu().numpy()))
    assert sum(int(m[2][2].sum()) for m in marks if m[0] == "step") > N  # many episodes ended
    launches = np.stack(log)  # [launches, N]
    final = env.game.batch.export()
    for i in range(0, N, 5):
        # the opponent decisions of table i, in order: every non-skipped launch that is not the agent's own step
        agent_launch = {m[1] for m in marks if m[0] == "step"}
        queue = [int(launches[k, i]) for k in range(len(launches)) if k not in agent_launch and launches[k, i] != SKIP]
        queue.reverse()
        ref = RefEnvOnOracle(ora, [lambda mask: queue.pop()] * (P - 1), env_id=i, **cfg)
        for kind, _, payload in marks:
            if kind == "reset":
                if payload is None or payload[i]:
                    ref.reset()
            else:
                a, reward, done, hand = payload
                r_ref, d_ref, h_ref = ref.step(int(a[i]))
                assert (bits(reward[i]), bool(done[i]), bool(hand[i])) == (bits(r_ref), d_ref, h_ref), i
        assert not queue, f"table {i}: {len(queue)} logged opponent decisions were never used"
        snap = ref.o.snapshot()
        assert [bits(x) for x in final["credits"][:, i]] == [bits(x) for x in snap["credits"]], i
        assert int(final["hand"][i]) == snap["hand"] and int(final["active"][i]) == snap["active"], i


def test_skip_action_leaves_envs_untouched(torch_cuda, ora):
    torch = torch_cuda
    from pokerl_b200 import GameBatch
    from pokerl_b200 import _native as nat
    N = 2048
    g = GameBatch(N, 4, 100, 2, 1, seed=4)
    o = ora.OracleBatch(N, 4, 100, 2, 1, seed=4)
    g.reset(); o.reset()
    for t in range(24):
        acts = g.sample_actions(t).clone()
        acts[t % 3::3] = nat.ACTION_SKIP
        flags, valid = g.step(acts)
        f_cpu, v_cpu = o.step(acts.cpu().numpy())
        assert np.array_equal(flags.cpu().numpy(), f_cpu) and np.array_equal(valid.cpu().numpy(), v_cpu), t
        assert ((flags.cpu().numpy() == nat.F_SKIPPED) == (acts.cpu().numpy() == nat.ACTION_SKIP)).all()
    ge, oe = g.export(), o.export()
    for k in ("credits", "bets", "pending", "payoffs", "minraise"):
        assert np.array_equal(np.ascontiguousarray(ge[k]).view(np.uint64), np.ascontiguousarray(oe[k]).view(np.uint64)), k
    st = g.stats()
    assert st["invalid"] == 0 and st["steps"] == 24 * N - sum(len(range(t % 3, N, 3)) for t in range(24))


def test_q_learning_example_runs(torch_cuda):
    """examples/q_learning.py (batched, config 5 shape) for a few steps on one GPU: the Q table moves, is finite,
    and the hand-ranking features are valid categories."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("q_learning_example", os.path.join(here, "examples", "q_learning.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Q, score, stats = mod.run(num_envs=2048, num_steps=12, quiet=True)
    assert Q.shape == (10, 7) and bool(torch_cuda.isfinite(Q).all()) and float(Q.abs().sum()) > 0
    assert score[1] > 0 and stats["invalid"] == 0 and stats["poisoned"] == 0 and stats["hands"] > 0


@pytest.mark.parametrize("P,weights,phases", [(3, None, False), (4, (10, 30, 30, 5, 5, 5, 45), False), (2, None, False),
                                              (3, None, True), (4, (10, 30, 30, 5, 5, 5, 45), True), (2, None, True),
                                              (9, None, True)])
def test_fused_env_step_and_reset_equal_the_reference_loops(torch_cuda, ora, P, weights, phases):
    """pkb_env_step / pkb_env_reset (ONE launch per PokerGameEnv.step / reset, opponents played by the device sampler
    inside the kernel) against game_env.py's loops restated over the oracle, table by table: rewards (bit for bit),
    done / hand_over, the valid mask handed to the agent and the final state.  The j-th opponent move of a table
    inside call c draws with sampler index c * ENV_ROUND_CAP + j - 1.  phases: PokerGameEnv.step as a sequence of
    launches, the tables whose opponents are still playing after 3 / 9 / 21 / 45 moves packed into the scratch arena
    (pkb_set_env_phases; large handles do that by default)."""
    torch = torch_cuda
    from pokerl_b200 import _native as nat
    from pokerl_b200.agents import RandomAgent
    from pokerl_b200.envs import PokerGameEnv
    N, T, seed = 640, 50, 13
    cfg = dict(num_players=P, start_credits=30, big_blind=4, small_blind=2, seed=seed)
    env = PokerGameEnv(agents=[RandomAgent(weights) for _ in range(P - 1)], num_envs=N, **cfg)
    assert env._fused
    env.game.batch.set_env_phases(0 if phases else -1)
    gen = torch.Generator(device="cuda").manual_seed(3)
    calls = []  # ("reset", mask or None) / ("step", actions, reward, done, hand, valid)
    state = env.reset()
    calls.append(("reset", None, state.valid_mask.cpu().numpy()))
    for t in range(T):
        assert bool((env.game.batch.active_player == 0).all())
        a = torch.multinomial(state.valid_actions, 1, generator=gen)[:, 0].to(torch.int32)
        state, reward, done, hand = env.step(a)
        calls.append(("step", a.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(), hand.cpu().numpy()))
        if bool(done.any()):
            state = env.reset(env_mask=done)
            calls.append(("reset", done.cpu().numpy(), state.valid_mask.cpu().numpy()))
    assert sum(int(c[3].sum()) for c in calls if c[0] == "step") > N  # many episodes ended
    final = env.game.batch.export()
    w = None if weights is None else list(weights)
    for i in range(0, N, 7):
        ctr = {"base": 0, "j": 0}

        def opponent(mask):
            a = ora.sample_action(seed, i, ctr["base"] + ctr["j"], mask, w)
            ctr["j"] += 1
            return a

        ref = RefEnvOnOracle(ora, [opponent] * (P - 1), env_id=i, **cfg)
        for c, call in enumerate(calls):
            ctr["base"], ctr["j"] = c * nat.ENV_ROUND_CAP, 0
            if call[0] == "reset":
                if call[1] is None or call[1][i]:
                    ref.reset()
                    assert ref.o.valid_mask() == int(call[2][i]), (i, c)
            else:
                _, a, reward, done, hand = call
                r_ref, d_ref, h_ref = ref.step(int(a[i]))
                assert (bits(reward[i]), bool(done[i]), bool(hand[i])) == (bits(r_ref), d_ref, h_ref), (i, c)
        snap = ref.o.snapshot()
        assert [bits(x) for x in final["credits"][:, i]] == [bits(x) for x in snap["credits"]], i
        assert [bits(x) for x in final["payoffs"][:, i]] == [bits(x) for x in snap["payoffs"]], i
        assert int(final["hand"][i]) == snap["hand"] and int(final["active"][i]) == snap["active"], i
        assert int(final["episode"][i]) == ref.o.g.episode, i


def test_fused_env_skips_and_rejects(torch_cuda, ora):
    torch = torch_cuda
    from pokerl_b200 import _native as nat
    from pokerl_b200.agents import RandomAgent
    from pokerl_b200.envs import PokerGameEnv
    N = 256
    env = PokerGameEnv(agents=[RandomAgent(), RandomAgent()], num_envs=N, num_players=3, start_credits=50, seed=2)
    env.reset()
    before = env.game.batch.export()
    acts = torch.full((N,), nat.ACTION_SKIP, dtype=torch.int32, device="cuda")
    reward, done, _ = env.game.batch.env_step(acts, 10 * nat.ENV_ROUND_CAP)
    assert bool((done == nat.F_SKIPPED).all()) and float(reward.abs().sum()) == 0.0
    after = env.game.batch.export()
    assert all(np.array_equal(before[k], after[k]) for k in before)
    with pytest.raises(ValueError):
        env.step(torch.full((N,), 1, dtype=torch.int32, device="cuda"))  # CHECK facing the blinds is invalid


def test_env_phases_at_size_equal_the_single_launch(torch_cuda):
    """PokerGameEnv.step on 2^17 four-player tables (thousands of stragglers per call, every phase populated) in phases
    -- the default at this size -- and as ONE launch: rewards, done / hand flags, valid masks and the whole state are
    identical after every call, resets of finished games included."""
    torch = torch_cuda
    from pokerl_b200.agents import RandomAgent
    from pokerl_b200.envs import PokerGameEnv
    N, T = (1 << 17) + 333, 14
    envs = []
    for phases in (True, False):
        env = PokerGameEnv(agents=[RandomAgent() for _ in range(3)], num_envs=N, num_players=4, start_credits=60, big_blind=4,
                           small_blind=2, seed=21, check_actions=False)
        env.game.batch.set_env_phases(0 if phases else -1)
        envs.append((env, env.reset()))
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(T):
        (ea, sa), (eb, sb) = envs
        assert torch.equal(sa.valid_mask, sb.valid_mask), t
        a = torch.multinomial(sa.valid_actions, 1, generator=gen)[:, 0].to(torch.int32)
        sa, ra, da, ha = ea.step(a)
        sb, rb, db, hb = eb.step(a)
        assert torch.equal(ra.view(torch.int64), rb.view(torch.int64)) and torch.equal(da, db) and torch.equal(ha, hb), t
        assert int(da.sum()) > 0 or t < 3
        xa, xb = ea.game.batch.export(), eb.game.batch.export()
        for k in xa:
            assert np.array_equal(np.ascontiguousarray(xa[k]).view(np.uint8), np.ascontiguousarray(xb[k]).view(np.uint8)), (t, k)
        if bool(da.any()):
            sa, sb = ea.reset(env_mask=da), eb.reset(env_mask=db)
        envs = [(ea, sa), (eb, sb)]
    st_a, st_b = ea.game.batch.stats(), eb.game.batch.stats()
    assert all(st_a[k] == st_b[k] for k in ("steps", "hands", "showdowns", "games", "invalid", "poisoned"))
