"""One training step (get_loss + backward) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import torch
import bench

def main(cfg="cc12m_64x64", B=64, mode="train"):
    dev = torch.device("cuda", 0)
    pipe, nested = bench.build_pipeline(cfg, dev)
    host = bench.synthetic_host_batch(cfg, B, 1234)
    sample = {k: v.to(dev) for k, v in host.items()}
    vm = pipe.get_model().vision_model
    def step():
        if mode == "train":
            pipe.train()
            loss, *_ = pipe.get_loss(sample)
            loss.mean().backward()
            vm.zero_grad(set_to_none=True)
        else:
            pipe.sample(B, sample, bench.RES[cfg][0], dev, num_inference_steps=2, ddim_eta=0.0, resample_steps=True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cc12m_64x64", int(sys.argv[2]) if len(sys.argv) > 2 else 64,
         sys.argv[3] if len(sys.argv) > 3 else "train")
