"""One-off parity fuzz on the GPU: generated populations (both mixes, several sizes) scored through the C ABI and
compared with the conditioned oracle; a mismatch counts only if some mismatching POINT is one where the oracle's own
value is stable under +-1 ulp jitter of its math library.
    python scripts/fuzz_parity.py [extreme|synth] [first_seed] [n_seeds]"""
import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import importlib.util
spec=importlib.util.spec_from_file_location("t","tests/test_gpu_parity.py"); t=importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
import closyr_b200 as cb
from oracle import oracle as O
ctx=cb.Context([0])
xs=np.array([-np.inf,-1e308,-1e154,-1e6,-3.0,-1.0,-0.5,-1e-300,-5e-324,-0.0,0.0,5e-324,1e-300,0.5,1.0,2.0,710.0,1e6,1e154,1e308,np.inf, 0.1,0.2,0.3,0.7,0.9,1.1,1.3,1.7,2.5,3.14159,4.0,5.5,7.0,9.0,10.0,-10.0,-7.0,-2.0,-0.1])
if len(sys.argv)>1 and sys.argv[1]=='synth':
    xs=(np.pi/15.0)*49.0*np.arange(256)/255.0
    ys=2.0+1.5*xs/np.pi+np.cos(xs)
else:
    ys=np.linspace(-2,2,len(xs))
first=int(sys.argv[2]) if len(sys.argv)>2 else 100
count=int(sys.argv[3]) if len(sys.argv)>3 else 40
bad=0; tot=0; chaotic=0
for seed in range(first,first+count):
    for mix,ml in ((0,40),(1,80),(0,12)):
        b=cb.PopulationBuilder(); texts=b.generate(seed,500,ml,mix,want_texts=True)
        ctx.set_points(xs,ys)
        got,_=ctx.score_batch(b.view(),ml)
        want,ok=O.score_batch_conditioned(texts,xs,ys,ml)
        err=np.abs(got-want)/np.maximum(1.0,np.abs(want))
        for i in np.nonzero(err>1e-9)[0]:
            if not O.score_discrepancy_explained(texts[i],xs,ys,ml,got[i],want[i]):
                # per point: is every mismatching point one where the oracle's own value moves under +-1 ulp jitter?
                bb=cb.PopulationBuilder(); bb.add(texts[i]); f,_=ctx.eval_batch(bb.view(),xs); f=f[0]
                base=O.eval_vec_pheno(texts[i],xs)
                fin=np.where(np.isfinite(f),f,np.inf)
                with np.errstate(invalid="ignore"):
                    todo=set(np.nonzero(~((fin==base)|(np.abs(fin-base)<=1e-9*np.abs(base))))[0].tolist())
                # two libraries that are each within 1.5 ulp can be 3 ulp apart, the jitter explores +-1 ulp around
                # the oracle's own library: a point is explained when the jitter reproduces a quarter of the deviation
                for sd in range(1,513):
                    if not todo: break
                    O.lib().oracle_set_jitter(sd); v=O.eval_vec_pheno(texts[i],xs)
                    for k in list(todo):
                        if np.isinf(fin[k])!=np.isinf(base[k]): hit=np.isinf(v[k])==np.isinf(fin[k])
                        else: hit=np.isinf(v[k])!=np.isinf(base[k]) or abs(v[k]-base[k])>=0.25*abs(fin[k]-base[k]) or abs(v[k]-base[k])>=0.1*abs(base[k])   # last: the value is effectively random here
                        if hit: todo.discard(k)
                O.lib().oracle_set_jitter(0)
                real_bad=np.array([len(todo)>0])
                detail=[(xs[k],fin[k],base[k]) for k in sorted(todo)]
                if real_bad.any():
                    bad+=1
                    if bad<=10: print("UNEXPLAINED",seed,mix,texts[i],got[i],want[i],detail)
                else: chaotic+=1
        tot+=len(texts)
print("expressions",tot,"unexplained",bad,"chaotic-only",chaotic)
