import numpy as np, torch, sys
sys.path.insert(0, '.')
import salsa_b200
from oracle import salsa as osalsa
g = np.load('tests/golden/clip_cases.npz')
audio = g['audio_foa']
ref = osalsa.multichannel_stft(audio, 512, 300)[1:256].astype(np.complex64)
out = salsa_b200.stft(audio, n_fft=512, hop_length=300, lower_bin=1, upper_bin=256)
neq = (out != ref)
print('mismatch frac', neq.mean())
print('per-frame mismatch (first 10, last 5):', neq.mean(axis=(0,2))[:10], neq.mean(axis=(0,2))[-5:])
print('per-bin mismatch (every 16th):', neq.mean(axis=(1,2))[::16])
print('per-chan', neq.mean(axis=(0,1)))
d = np.abs(out-ref); 
rel = d/np.maximum(np.abs(ref),1e-30)
print('max rel', rel.max(), 'median rel among mismatches', np.median(rel[neq]))
mag = np.abs(ref)
for lo,hi in [(0,1e-6),(1e-6,1e-5),(1e-5,1e-4),(1e-4,1e-3),(1e-3,1e-2),(1e-2,1e-1),(1e-1,1),(1,100)]:
    m = (mag>=lo)&(mag<hi)
    if m.sum(): print('mag [%g,%g): n=%d mismatch=%.4f' % (lo,hi,m.sum(),neq[m].mean()))
