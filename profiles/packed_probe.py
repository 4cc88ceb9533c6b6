"""Device time per step of pkb_step_packed with the actions already on the device (no copies): six handles on six
streams, the kernel side of the bench's end-to-end leg."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pokerl_b200 import GameBatch  # noqa: E402

B, R, N = int(os.environ.get("PKB_PROBE_BATCHES", "6")), 24, 1 << 20
main = torch.cuda.current_stream()
streams = [torch.cuda.Stream() for _ in range(B)]
acts, runs = [], []
for j in range(B):
    rec = GameBatch(N, 2, 100, 2, 1, seed=0, env_id_base=j * N, auto_reset=True)
    rec.reset()
    for r in range(64):
        rec.step_random(r)
    a = []
    for r in range(R):
        rec.step_random(64 + r, record_actions=True)
        a.append(GameBatch.pack_actions(rec.actions_taken))
    acts.append(a)
    run = GameBatch(N, 2, 100, 2, 1, seed=0, env_id_base=j * N, auto_reset=True)
    run.reset()
    for r in range(64):
        run.step_random(r)
    runs.append(run)
outs = [torch.empty(N, dtype=torch.uint8, device="cuda") for _ in range(B)]
for j in range(B):
    with torch.cuda.stream(streams[j]):
        streams[j].wait_stream(main)
        runs[j].refresh_valid()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for st in streams:
    st.wait_stream(main)
for r in range(R):
    for j in range(B):
        with torch.cuda.stream(streams[j]):
            runs[j].step_packed(acts[j][r], out=outs[j])
for st in streams:
    main.wait_stream(st)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) * 1e3 / (R * B)
print(f"step_packed, {B} handles on {B} streams, actions resident: {us:.1f} us/step, {N / us / 1e3:.2f} G env-steps/s; invalid {sum(x.stats()['invalid'] for x in runs)}")
