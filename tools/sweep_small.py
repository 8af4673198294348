import sys, json, torch, warnings
sys.path.insert(0,'/root/repo'); warnings.simplefilter('ignore')
from salib_b200 import _runtime as rt
from salib_b200.sample import sobol as sm
def t(fn, steps=5, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/1e3/steps
for D,l2 in [(16,22),(32,20),(64,20),(128,18),(8,22)]:
    N=1<<l2; p={"num_vars":D,"names":[str(i) for i in range(D)],"bounds":[[0.,1.]]*D}
    plan=sm.SamplePlan(p,True,sm.direction_vectors(2*D),None); step=2*D+2
    buf=rt.empty((step*N,D),torch.float64)
    s=t(lambda: plan.generate(0,N,out=buf)); print(D,l2,round(step*N*D*8/s/1e9),"GB/s")
    del buf
