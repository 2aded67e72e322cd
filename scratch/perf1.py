import sys, time
sys.path.insert(0, 'tests'); sys.path.insert(0, 'aero-cli_b200')
import numpy as np
import aeroddc
nv = int(sys.argv[1]) if len(sys.argv)>1 else 256
Fs=61440000; B=Fs//4; D=8; L=5
t0=time.time()
bank = aeroddc.Bank(Fs,B,aeroddc.CF32,0)
rng = np.random.default_rng(1)
freqs = rng.integers(int(-0.45*Fs), int(0.45*Fs), nv)
for i,f in enumerate(freqs): bank.add_vfo(float(f),D,L,0,0.05,1,1,1,"V%04d"%i)
import os
if os.environ.get("AERODDC_FAST"): bank.set_mode(aeroddc.MODE_FAST)
bank.finalize()
print('finalize %.2fs, dev MB %.1f'%(time.time()-t0, bank.device_bytes()/1e6))
s0 = bank.host_slot(0); s1 = bank.host_slot(1)
s0[:] = (rng.standard_normal(2*B)*0.1).astype(np.float32); s1[:] = s0[::-1]
for it in range(6):
    t1=time.time()
    bank.submit(s0 if it%2==0 else s1); bank.wait()
    ms, n = bank.last_timing(); mm = bank.last_main_ms()
    print('blk',it,'wall %.1f ms'%((time.time()-t1)*1e3),'kern %.3f ms main %.3f ms launches %d'%(ms,mm,n),'Gsps(main) %.1f  Gsps(kern) %.1f'%(nv*B/mm/1e6, nv*B/ms/1e6))
# pipelined
t1=time.time(); K=8
bank.submit(s0)
for it in range(K):
    bank.submit(s1 if it%2==0 else s0); bank.wait()
bank.wait()
dt=time.time()-t1
print('pipelined e2e: %.1f Gsps'%(nv*B*(K+1)/dt/1e9))
