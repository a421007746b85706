"""Randomised check of Env.reset / set_seed chains: the unmodified reference (one process per chain) against the host topology
loader (seed bookkeeping: next_topology_seed, set_seed) + the C oracle.  Build-container tool (needs oracle/_ref).

    python tools/fuzz_cim_reset_chain.py [rng_seed] [n_chains]
"""
import sys, os, json, multiprocessing as mp, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

def ref_run(spec, q):
    os.environ["SKIP_DEPLOYMENT"]="TRUE"
    sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref')); sys.path.insert(1, os.path.join(ROOT, 'oracle', '_ref', '_stubs'))
    from maro.simulator import Env
    env=Env("cim",spec["topology"],durations=spec["durations"])
    out=[]
    for op in spec["ops"]:
        if op[0]=="seed": env.set_seed(op[1]); env.reset(keep_seed=True)
        elif op[0]=="reset": env.reset(keep_seed=op[1])
        m,d,done=env.step(None); n=0; first=None
        while not done:
            if first is None: first=[int(x) for x in (d.tick,d.port_idx,d.vessel_idx,d.action_scope.load,d.action_scope.discharge)]
            m,d,done=env.step(None); n+=1
        out.append([n,first,[int(m["order_requirements"]),int(m["container_shortage"])]])
    q.put(out)

def ours(spec):
    from maro_b200.scenarios.cim.topology import build_topology, next_topology_seed, load_config
    from oracle.cim_oracle import CimOracle
    conf=load_config(spec["topology"]); mt=spec["durations"]
    topo=build_topology(conf,mt); pending=None; out=[]
    for op in spec["ops"]:
        if op[0]=="seed": topo=build_topology(conf,mt,seed=op[1])
        elif op[0]=="reset":
            if not op[1]: topo=build_topology(conf,mt,seed=next_topology_seed(topo))
            # keep_seed=True: same topology
        o=CimOracle(topo); st,d,m=o.step(None); n=0; first=None
        while st==0:
            if first is None: first=[int(x) for x in d[:5]]
            st,d,m=o.step(None); n+=1
        out.append([n,first,[int(m[0]),int(m[1])]])
    return out

def main():
    rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
    tops=["toy.4p_ssdd_l0.%d"%k for k in (0,2,5,8)]+["toy.5p_ssddd_l0.3","toy.6p_sssbdd_l0.6","global_trade.22p_l0.4"]
    bad=0
    for case in range(int(sys.argv[2]) if len(sys.argv)>2 else 12):
        t=str(rng.choice(tops)); ops=[("run",)]
        for k in range(int(rng.integers(2,5))):
            r=rng.random()
            ops.append(("seed",int(rng.integers(1,50000))) if r<0.3 else ("reset",bool(r<0.5)))
        spec=dict(topology=t,durations=int(rng.integers(40,60 if t.startswith("global") else 120)),ops=ops)
        ctx=mp.get_context("spawn"); q=ctx.Queue(); p=ctx.Process(target=ref_run,args=(spec,q)); p.start(); ref=q.get(); p.join()
        got=ours(spec)
        ok=json.dumps(ref,default=int)==json.dumps(got,default=int)
        bad+=not ok
        print(case,"ok" if ok else "MISMATCH",spec["topology"],ops, "" if ok else (ref,got),flush=True)
    print("mismatches:",bad)
if __name__=="__main__": main()
