import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rosdyn_b200 import fixtures, _lib
if '--lib' in sys.argv:
    k = sys.argv.index('--lib'); _lib.set_library_path(os.path.abspath(sys.argv[k + 1])); del sys.argv[k:k + 2]
from rosdyn_b200.chain import Chain, fill_uniform
name = sys.argv[1] if len(sys.argv) > 1 else "c6"
d = fixtures.by_name(name); ch = Chain(d)
names = ch.getActiveJointsName()
ch.setComponents([{"type": "friction1", "joint": n, "min_velocity": 0.01, "max_velocity": 2.0} for n in names])
S = 8_000_000
q, dq, ddq = (fill_uniform(d.n_inputs, S, 1, s, device="cuda") for s in range(3))
for f, nm in ((ch.regressorGramExt, "ext"), (ch.regressorGram, "rigid")):
    for _ in range(2): f(q, dq, ddq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f(q, dq, ddq)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(name, nm, ms, "ms", S / ms / 1e6, "G samples/s", flush=True)
