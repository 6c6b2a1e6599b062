"""N training steps (B=128, bf16, RGBNT201 config) and nothing else -- the target of the ncu runs of
tools/profile_round.sh (bench.py also measures e2e, eval forward, evaluation metrics ...: minutes under ncu).
Usage: python tools/one_step.py [steps=4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device("cuda", 0)
    from editor_b200.train import Trainer
    model, sd, x, label, cam, _ = bench.build_case(dev, 128, seed=1)
    model.train()
    tr = Trainer(model)
    xg = {k: v.to(dev) for k, v in x.items()}
    lg, cg = label.to(dev), cam.to(dev)
    for _ in range(steps):
        tr.step(xg, lg, cg)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
