import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'aero-cli_b200')
import numpy as np, aeroddc
from case_util import ALL_CASES, case_fmt, raw_block, float_block, parity_metrics
from oracle_bind import Oracle
for d in ALL_CASES:
    if not d["demod_usb"]: continue
    bank = aeroddc.Bank(d["Fs"], d["B"], case_fmt(d), 0)
    bank.add_vfo(d["mixer"], d["D"], d["L"], d["filter_bw"], d["gain"], 1, 1, 1, "X")
    bank.set_mode(aeroddc.MODE_FAST)
    bank.finalize()
    o = Oracle(d["Fs"], d["B"], d["D"], d["L"], d["mixer"], d["gain"], d["filter_bw"])
    worst=(0,1e9); srel=0; rms=0
    for b in range(d["blocks"]):
        bank.process(raw_block(d,b)); got,_=bank.output(0); want=o.process(float_block(d,b))
        g=np.frombuffer(got,np.int16); w=np.frombuffer(want,np.int16)
        m=parity_metrics(g,w); worst=(max(worst[0],m[0]),min(worst[1],m[1]))
        sg=bank.stage_d(0,d["B"]>>d["D"]); so=o.stage(d["D"])
        srel=max(srel, float(np.abs(sg-so).max()/max(np.abs(so).max(),1e-30)))
        rms=max(rms,float(np.sqrt((w.astype(float)**2).mean())))
    print('%-24s max|err|/FS %.2e (%.1f LSB)  SNR %.1f dB  stageD rel err %.2e  out rms %.0f LSB' % (d["name"],worst[0],worst[0]*32768,worst[1],srel,rms))
    bank.close()
