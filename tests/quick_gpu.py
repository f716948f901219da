import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, synth, vacmap_b200 as vb
n = int(sys.argv[1]) if len(sys.argv)>1 else 2000
passes = int(sys.argv[2]) if len(sys.argv)>2 else 3
import os
workers = int(os.environ.get('VM_QUICK_WORKERS', '0'))
t0=time.time(); ref = synth.make_reference(1, 5_000_000); print('ref', time.time()-t0)
t0=time.time(); reads = synth.make_reads(ref, 11, n, read_len=15000, err=0.10); print('reads', time.time()-t0)
t0=time.time(); ix = vb.Index(ref); print('index', time.time()-t0, ix.n_keys, ix.n_minimizers, ix.mid_occ)
al = vb.Aligner(ix, vb.default_option('H'), 'H', workers=workers)
enc = [s.upper().encode() for _, s in reads]
off = np.zeros(len(reads)+1, dtype=np.int64)
for i,e in enumerate(enc): off[i+1]=off[i]+len(e)
cat = b"".join(enc)
for it in range(passes):
    t0=time.time(); rec_off, recs, cig = al.align_packed(cat, off); dt=time.time()-t0
    mapped = int((np.diff(rec_off)>0).sum())
    print('iter', it, 'sec', round(dt,3), 'reads/s', round(n/dt,1), 'Gbp/s', round(off[-1]/dt/1e9,4), 'mapped', mapped, 'records', len(recs))
    print('  ', {k: round(v,1) for k,v in sorted(al.last_stage_ms.items())})
