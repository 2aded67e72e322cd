import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, 'aero-cli_b200')
import numpy as np, aeroddc
from oracle_bind import Oracle, synth_anchor
Fs,B,D,L,f,g,bw = 1536000,384000,5,0,123456.0,0.5,0
bank = aeroddc.Bank(Fs,B,aeroddc.CF32,0); bank.add_vfo(f,D,L,bw,g,1,1,1,"V"); bank.finalize()
o = Oracle(Fs,B,D,L,f,g,bw)
for b in range(3):
    x = synth_anchor(b*B,B)
    try:
        bank.process(x)
    except Exception as e:
        print('ERR', str(e)[:200]); sys.exit(1)
    print('blk',b, bank.output(0)[0]==o.process(x))
