import sys, time, types, importlib
sys.path.insert(0, '/root/repo')
import numpy as np
import asr_b200 as A
from oracle import speechpy_ref as R, sox_ref as SX
def chk(name, got, want):
    worst=0
    for a,b in zip(got,want):
        assert a.shape==b.shape,(name,a.shape,b.shape)
        if a.size==0: continue
        e=np.abs(a.astype(np.float64)-b.astype(np.float64)); worst=max(worst,e.max())
    print(name,'max err',worst, flush=True)
pcm = A.synth.corpus(20, 1.0, 6.0, seed=3) + [np.zeros(1000,np.int16), A.synth.corpus(1,1,2,5)[0][:560], A.synth.corpus(1,1,2,5)[0][:470]]
for ft,D in (('mfcc',13),('fbank',80),('fbank',40)):
    for cm in (True,False):
        args = types.SimpleNamespace(frame_step=10, frame_length=25, feat_dim=D, feat_type=ft, cmvn=cm)
        f,l = A.process_pcm(pcm,args); w,wl = R.process_audios(pcm,args); assert l==wl,(l,wl)
        if ft=='fbank':
            # relative error for linear energies
            chk('%s D=%d cmvn=%s'%(ft,D,cm), [x/(np.abs(y)+1e-30) if not cm else x for x,y in zip(f,w)], [y/(np.abs(y)+1e-30) if not cm else y for y in w])
        else: chk('%s D=%d cmvn=%s'%(ft,D,cm), f, w)
args = types.SimpleNamespace(frame_step=10, frame_length=25, feat_dim=13, feat_type='mfcc', cmvn=True)
f,l = A.process_pcm(pcm[:20],args,delta_mode='time_regression'); w,_=R.process_audios(pcm[:20],args,delta_mode='time_regression'); chk('time_regression',f,w)
f,l = A.process_pcm(pcm[:20],args,bin_map='nfft_plus_one'); w,_=R.process_audios(pcm[:20],args,bin_map='nfft_plus_one'); chk('full spectrum',f,w)
f,l = A.process_pcm(pcm[:20],args,preemph=0.98); w,_=R.process_audios(pcm[:20],args,preemph=0.98); chk('preemph',f,w)
win=np.hamming(400); f,l = A.process_pcm(pcm[:20],args,window=win); w,_=R.process_audios(pcm[:20],args,window=win); chk('hamming',f,w)
fe = A.Frontend(A.FrontendConfig())
for s in (0.9,1.1):
    y = fe.perturb(pcm[:5], speeds=[s]*5); ref=[SX.speed_perturb(p,s) for p in pcm[:5]]
    d=[np.abs(a.astype(int)-b.astype(int)) for a,b in zip(y,ref)]; print('speed',s,[len(a) for a in y]==[len(b) for b in ref], max(x.max() for x in d), np.mean([ (x>0).mean() for x in d]))
    f = fe.extract(pcm[:5], speeds=[s]*5); w=[R.features_one(r) for r in ref]; chk('speed feats %s'%s, f, w)
y = fe.perturb(pcm[:5], gains=[0.8,1.5,1.23,1.0,2.5]); ref=[SX.volume_perturb(p,g) for p,g in zip(pcm[:5],[0.8,1.5,1.23,1.0,2.5])]
print('gain', max(np.abs(a.astype(int)-b.astype(int)).max() for a,b in zip(y,ref)))
f = fe.extract(pcm[:5], gains=[0.8,1.5,1.23,1.0,2.5]); chk('gain feats', f, [R.features_one(r) for r in ref])
# timing, device resident
import torch
lens = A.synth.durations(4000, 2, 15, np.random.default_rng(1))
pk_off = np.concatenate(([0], np.cumsum((lens+7)//8*8)))[:-1]
tot = int(((lens+7)//8*8).sum())
d = (torch.randn(tot, device='cuda')*3000).to(torch.int16)
fe.set_profiling(True)
out=None
for it in range(4):
    torch.cuda.synchronize(); t=time.time()
    out, oo, nfr = fe.run_packed(d, pk_off, lens, out=out)
    fe.sync(); torch.cuda.synchronize(); dt=time.time()-t
    print('iter',it,'wall ms',dt*1e3, fe.kernel_ms(), 'audio-h', lens.sum()/16000/3600, 'audio-h/s (k1+k2)', lens.sum()/16000/3600/((fe.kernel_ms()['device_pass'])*1e-3), flush=True)
