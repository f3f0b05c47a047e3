"""TEST INFRASTRUCTURE (CPU only).  Where does the agreement between ANY implementation and the reference stop?

The reference rotates every sample by exp(-i fl32(PHASE + fl32(w * t_n))) with float32 t_n (gpslib.py:1053-1054,
1343-1346): at 5 kHz x 32 ms the argument reaches 1000 rad, where one float32 ulp is 6e-5 rad, so every sample carries
its own pseudo-random phase error of that size.  An implementation that does not evaluate one float32 sin/cos argument
per sample (the CUDA tracker factorises the NCO: 2 sincos per thread and epoch instead of 16 N_CYC) computes the
mathematically exact rotation instead.  This script runs the oracle (bit-exact restatement of SatStream, pinned by
tests/test_oracle_golden.py) against ITSELF with only that one change -- everything else stays numpy float64/complex128
-- on the golden scenarios and prints how far FREQ / PHASE / AMPLITUDE / STD_DEV / prompts move.  Those numbers are the
floor under the tolerances of tests/test_track_gpu.py (output committed as profiles/parity_floor_r02.txt).

    python oracle/parity_floor.py
"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gps_oracle as orc
from gps_sdr_receiver_b200 import synth
import hashlib
F32=np.float32
def wipe_exact(data, freq, phase, n, t):
    w = 2*np.pi*freq          # same scalar as the reference (float32 if freq is float32)
    w64 = float(w); p64=float(phase)
    nn = np.arange(1,n+1,dtype=np.float64)/orc.SAMPLE_RATE
    rot = np.exp(-1j*(p64 + w64*nn))
    out = (rot*data[:n].astype(np.complex128)).astype(np.complex64)
    ph = phase + 2*np.pi*freq*t[n-1]       # the carried phase as the reference computes it
    return out, np.remainder(ph, 2*np.pi)
for n_cyc in (32,8):
    g=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests','golden',f'traj_ncyc{n_cyc}.npz'))
    sats=synth.default_constellation(6,seed=5)
    nE=int(g['n_epochs']); raw=synth.make_iq(sats,n_cyc*nE,noise_sigma=0.25,seed=11)
    ngps=n_cyc*2048; s0=int(g['start_epoch'])
    tot=eq=0; maxulp=0; maxph=0; maxamp=0; maxstd=0; maxpr=0
    for ci,(p,f,d) in enumerate(g['chan_init']):
        a=orc.Channel(int(p),float(f),delay=int(d),n_cyc=n_cyc)
        b=orc.Channel(int(p),float(f),delay=int(d),n_cyc=n_cyc)
        for e in range(s0,nE):
            if e==int(g['gap_at']): continue
            blk=orc.raw_to_complex(raw[e*2*ngps:(e+1)*2*ngps]); smp=np.int64((e+1)*ngps)
            a.process(blk,smp)
            old=orc.wipeoff; orc.wipeoff=wipe_exact
            try: b.process(blk,smp)
            finally: orc.wipeoff=old
            fa,fb=F32(a.freq),F32(b.freq)
            tot+=1; eq+= (fa==fb)
            u=abs(int(fa.view(np.int32))-int(fb.view(np.int32))); maxulp=max(maxulp,u)
            dph=abs((float(a.phase)-float(b.phase)+np.pi)%(2*np.pi)-np.pi); maxph=max(maxph,dph)
            if a.amplitude and not a.sweep and a.locked==b.locked:
                maxamp=max(maxamp,abs(float(a.amplitude)/float(b.amplitude)-1)); maxstd=max(maxstd,abs(float(a.std_dev)/float(b.std_dev)-1))
                if len(a.prompt)==len(b.prompt):
                    maxpr=max(maxpr,np.abs(a.prompt-b.prompt).max()/np.abs(a.prompt).max())
    print(n_cyc,"epochs",tot,"FREQ bit-equal",eq/tot,"max ulp",maxulp,"max dPHASE",maxph,"amp rel",maxamp,"std rel",maxstd,"prompt complex rel",maxpr)
