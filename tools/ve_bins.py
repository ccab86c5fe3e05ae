import sys, json, math
sys.path.insert(0,'/root/repo')
from bnpp_b200 import capi, model, synth
import torch
ctx = capi.Context(0)
_, bn = model.from_uai_text(ctx, synth.random_bn_uai(64,40,4,5))
order, width = bn.order(list(range(64)), {}, "mf")
plan = bn.plan([], order)
res = torch.zeros(2, dtype=torch.float64, device="cuda")
for _ in range(3): plan.run(bn.table_ptrs, [], res.data_ptr(), res.data_ptr()+8)
ctx.sync()
# graph-replayed, no per-step events
s = ctx.torch_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(10): plan.run(bn.table_ptrs, [], res.data_ptr(), res.data_ptr()+8)
e1.record(s); e1.synchronize()
print("graph replay ms/query", e0.elapsed_time(e1)/10)
plan.set_profiling(True)
acc=None
for _ in range(3):
    plan.run(bn.table_ptrs, [], res.data_ptr(), res.data_ptr()+8)
    st = plan.step_stats()
    acc = st if acc is None else [dict(a, ms=a["ms"]+b["ms"]) for a,b in zip(acc,st)]
tot=sum(x["ms"] for x in acc)/3
print("event-per-step total", tot)
bins={}
for x in acc:
    b = int(math.log2(max(1,x["entries"])))//4*4
    d = bins.setdefault(b,[0,0.0,0])
    d[0]+=1; d[1]+=x["ms"]/3; d[2]+=x["bytes"]
for b in sorted(bins): print("entries 2^%d..2^%d: n=%d ms=%.3f GB=%.3f ideal_ms=%.3f"%(b,b+3,bins[b][0],bins[b][1],bins[b][2]/1e9,bins[b][2]/6.54e9/1e0*1e-0/1e3*1e3/1e0 if False else bins[b][2]/6.54e12*1e3))
