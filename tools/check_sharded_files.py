"""Multi-GPU check of the file-level drivers (run under torchrun, one rank per GPU):
process_libri_feats_sharded writes the reference's chunk layout with every rank extracting on its own GPU
(FLAC decode on the device); rank 0 compares the files with a single-GPU run of the same list."""
import os, sys, types, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist, joblib
import asr_b200 as A
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
root = "/dev/shm/asr_sharded_check"
if rank == 0:
    shutil.rmtree(root, ignore_errors=True); os.makedirs(root)
    pcm = A.synth.corpus(240, 1.0, 8.0, seed=808)
    packed, off, lens = A.pack_pcm(pcm)
    A.audio_io.write_audio_batch([os.path.join(root, "%04d.flac" % i) for i in range(len(pcm))], packed, off, lens, 16000)
dist.barrier()
paths = [os.path.join(root, "%04d.flac" % i) for i in range(240)]
args = types.SimpleNamespace(frame_step=10, frame_length=25, feat_dim=13, feat_type="mfcc", cmvn=True, feat_dir=os.path.join(root, "features"))
ok = True
for threshold, k, cat in ((10 ** 6, 1, "dev"), (100, 5, "train-100")):
    flen = A.sharding.process_libri_feats_sharded(paths, cat, k, args, threshold=threshold, device_decode=True)
    dist.barrier()
    if rank == 0:
        want, want_len = A.process_audios(paths, args, device=0)
        if k == 1:
            back = list(joblib.load(args.feat_dir + "/dev-feats.pkl"))
        else:
            back = [c for i in range(k) for c in joblib.load(args.feat_dir + "/%s-feats-%d.pkl" % (cat, i))]
        same = flen == want_len == np.load(args.feat_dir + "/%s-featlen.npy" % cat).tolist() and \
            all(np.array_equal(a, b) for a, b in zip(back, want)) and len(back) == 240
        print("%s: k=%d world=%d files identical to the single-GPU run: %s" % (cat, k, world, same), flush=True)
        ok = ok and same
dist.barrier()
if rank == 0:
    shutil.rmtree(root, ignore_errors=True)
    print("SHARDED FILE CHECK", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
