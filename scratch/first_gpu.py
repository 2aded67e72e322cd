import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, 'aero-cli_b200')
import numpy as np
from oracle_bind import Oracle, synth_anchor, fnv1a64
import aeroddc
print('fp32 peak', aeroddc.measure_fp32_peak(0))
cases = [(1536000,384000,5,0,123456.0,0.5,0,6,'4435d1a5843eed39'),(288000,57600,1,6,-34567.0,0.5,0,8,'78e7d3dd6abf468a'),
         (288000,57600,0,6,20000.0,0.25,3000,8,'4d86251b7e306418'),(1920000,480000,3,5,-250000.0,0.5,0,6,'c315edd09d50c16e')]
for Fs,B,D,L,f,g,bw,nb,want in cases:
    t0=time.time()
    bank = aeroddc.Bank(Fs,B,aeroddc.CF32,0)
    bank.add_vfo(f,D,L,bw,g,1,1,1,"V0001")
    bank.finalize()
    o = Oracle(Fs,B,D,L,f,g,bw)
    allb=b''; ok=True
    for b in range(nb):
        x = synth_anchor(b*B,B)
        bank.process(x)
        pg,rate = bank.output(0)
        po = o.process(x)
        sg = bank.stage_d(0, B>>D); so = o.stage(D)
        if not np.array_equal(sg, so):
            bad = np.nonzero(sg!=so)[0]
            print('  blk',b,'stageD mismatch', len(bad), 'first', bad[:6], 'maxerr', np.abs(sg-so).max())
        if pg!=po:
            a=np.frombuffer(pg,np.int16); c=np.frombuffer(po,np.int16); bad=np.nonzero(a!=c)[0]
            print('  blk',b,'payload mismatch',len(bad),'first',bad[:6], 'maxerr', np.abs(a.astype(int)-c).max()); ok=False
        allb+=pg
    print(Fs,D,L,'byte-identical' if ok else 'MISMATCH','%016x'%fnv1a64(allb), want, 'ms/blk kern', bank.last_timing(), 'wall %.1fs'%(time.time()-t0))
    bank.close()
