#!/usr/bin/env python
"""Randomised differential check on the CPU SIMT emulator: host vs device plan builder x half-warp vs 32-lane wavefronts, random
tables, references, query widths 1..256 and gap models (incl. zero penalties), every pair also against the oracle.
    python tools/fuzz_emu.py [seed]"""
import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'micall-lite_b200'), os.path.join(ROOT, 'tests', 'simt_emu')]
import numpy as np, build_emu
from gotoh_b200 import _ffi, packing
from gotoh_b200.api import Aligner
from oracle.oracle import Oracle
al = Aligner(_ffi.Library(build_emu.build())); ora = Oracle("port")
rng = random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
bad=0; total=0
for trial in range(14):
    matrix = rng.choice([0,0,1])
    alpha = "ACGTNRY-acgt" if matrix==0 else "ARNDCQEGHILKMFPSTWYVBZX*-"
    nrefs = rng.randint(1,5)
    refs = ["".join(rng.choice(alpha[:4] if matrix==0 and rng.random()<0.7 else alpha) for _ in range(rng.randint(1,260))) for _ in range(nrefs)]
    qs=[]; ridx=[]
    for k in range(220):
        r=rng.randrange(nrefs); a=refs[r]
        if rng.random()<0.6:
            lo=rng.randrange(len(a)); q=list(a[lo:lo+rng.randint(1,200)] or "A")
            for _ in range(rng.randint(0,4)): q[rng.randrange(len(q))]=rng.choice(alpha)
            q="".join(q)
        else: q="".join(rng.choice(alpha) for _ in range(rng.randint(1,256)))
        qs.append(q[:256]); ridx.append(r)
    gip=rng.choice([0,0,1,3,10,15,40]); gep=rng.choice([0,0,1,3,10]); term=rng.choice([0,1])
    rb,ro=packing.pack(refs); qb,qo=packing.pack(qs); ri=np.asarray(ridx,np.int32)
    res={}
    for prep in ("0","1"):
        os.environ["GOTOH_B200_DEVICE_PREP"]=prep
        for half in ("1","0"):
            os.environ["GOTOH_B200_HALF"]=half
            res[(prep,half)]=al.align_packed(rb,ro,ri,qb,qo,gip,gep,term,matrix)
    base=res[("0","1")]
    for key,v in res.items():
        for x in (0,1,3,4):
            if not (v[x]==base[x]).all(): bad+=1; print("MISMATCH builder/half",key,trial,x,gip,gep,term,matrix)
    fn=ora.align_it if matrix==0 else ora.align_it_aa
    for k in range(len(qs)):
        o,l=int(base[2][k]),int(base[3][k])
        got=(base[0][o:o+l].tobytes().decode('latin-1'),base[1][o:o+l].tobytes().decode('latin-1'),int(base[4][k]))
        exp=fn(refs[ridx[k]],qs[k],gip,gep,term)
        total+=1
        if got!=exp: bad+=1; print("ORACLE MISMATCH",trial,k,gip,gep,term,matrix,repr(refs[ridx[k]][:30]),repr(qs[k][:30]))
print("fuzz done: %d pairs, %d bad"%(total,bad))
