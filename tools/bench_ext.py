import sys, torch
sys.path.insert(0, '/root/repo')
from rosdyn_b200 import fixtures
from rosdyn_b200.chain import Chain, fill_uniform
d = fixtures.by_name("c6"); ch = Chain(d)
names = ch.getActiveJointsName()
ch.setComponents([{"type": "friction1", "joint": n, "min_velocity": 0.01, "max_velocity": 2.0} for n in names])
S = 4_000_000
q, dq, ddq = (fill_uniform(6, S, 1, s, device="cuda") for s in range(3))
for f, nm in ((ch.regressorGramExt, "ext"), (ch.regressorGram, "rigid")):
    for _ in range(2): f(q, dq, ddq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f(q, dq, ddq)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(nm, ms, "ms", S / ms / 1e6, "G samples/s")
