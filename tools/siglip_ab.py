"""Forward-sweep timings (SigLIP and ClipLoss, N = 32768, D = 512) in fresh processes, twice; the
environment knobs it sets are only read by builds that carry A/B variants of a kernel."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        import torch
        from latteclip_b200 import _lib
        import bench
        dev = torch.device("cuda:0")
        i, t = bench.synth_shard(32768, 512, 0, 1, set_id=9)
        ib, tb = i.to(dev).bfloat16(), t.to(dev).bfloat16()
        s, b = torch.tensor(10.0, device=dev), torch.tensor(-10.0, device=dev)
        for _ in range(5):
            loss = _lib.siglip_fwd(ib, tb, 0, s, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            loss = _lib.siglip_fwd(ib, tb, 0, s, b)
        e1.record()
        torch.cuda.synchronize()
        print(sys.argv[1], "siglip fwd ms", e0.elapsed_time(e1) / 30, "loss", float(loss))
        sc = torch.tensor(100.0, device=dev)
        for _ in range(5):
            out = _lib.clip_fwd(ib, tb, ib, tb, 0, sc)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(30):
            out = _lib.clip_fwd(ib, tb, ib, tb, 0, sc)
        e1.record()
        torch.cuda.synchronize()
        print(sys.argv[1], "clip fwd ms", e0.elapsed_time(e1) / 30, "loss", float(out[2]), "lse sum",
              float(out[0].double().sum()), float(out[1].double().sum()))
        print(sys.argv[1], bench.siglip_times(dev, reps=5))
    else:
        for rep in range(2):
            for v in ("0",):
                env = dict(os.environ, LATTE_B200_SIG_PACKED=v, LATTE_B200_FWD_PACKED=v)
                subprocess.run([sys.executable, os.path.abspath(__file__), "packed=" + v], env=env, check=True)
