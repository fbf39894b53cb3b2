import sys, os, torch, json
sys.path.insert(0, '/root/repo')
from spinn_b200 import synthetic
from spinn_b200.fused_step import PoissonInteriorStep
def time_ms(fn, reps=4):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for nodes, bn, samples in [(100, None, 1048576), (400, None, 1048576), (1024, 32, 1048576), (3600, 124, 1048576)]:
    pde, nn = synthetic.make_state(nodes=nodes, b_nodes=bn, samples=samples, device="cuda")
    xs, ys = (t.detach() for t in pde.p_samples)
    N, M = xs.numel(), nn.layer1.n
    for structured in (True, False):
        st = PoissonInteriorStep(nn, xs, ys, structured=structured)
        st.run(all_reduce=False)
        ms = time_ms(lambda: st.run(all_reduce=False))
        print(json.dumps({"N": N, "M": M, "structured": structured, "ms": ms, "evals_per_s": N*M/ms*1e3}))
