"""ViViT-B fact_encoder 16x224 batch 8 fwd+bwd timing (BASELINE config 3) + HOG kernel throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import ViViT, ClassificationHead
from videotransformer_pytorch_b200.graph import GraphedTrainStep
from videotransformer_pytorch_b200.hog import hog_features
dev = torch.device('cuda')

class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.m = ViViT(num_frames=16, img_size=224, patch_size=16, embed_dims=768, num_heads=12, num_transformer_layers=12)
        self.h = ClassificationHead(400, 768)
    def forward(self, x, y):
        return torch.nn.functional.cross_entropy(self.h(self.m(x)), y)

torch.manual_seed(0)
net = Net().to(dev).train()
x = torch.randn(8, 16, 3, 224, 224, device=dev); y = torch.randint(0, 400, (8,), device=dev)
step = GraphedTrainStep(net, (x, y))
for _ in range(3): step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step(x, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'ViViT-B fact_encoder 16x224 batch 8 fwd+bwd (graph): {ms:.2f} ms/step = {8 / ms * 1e3:.1f} clips/s; '
      f'{0.850e12 * 8 / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic')
fr = torch.randint(0, 256, (48, 224, 224, 3), dtype=torch.uint8, device=dev)
for _ in range(3): hog_features(fr)
torch.cuda.synchronize()
e0.record()
for _ in range(20): hog_features(fr)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
byt = 48 * (150528 + 84672)
print(f'HOG: 48 frames (16 samples x 3 centre frames) in {us:.1f} us = {48 / us * 1e6:.0f} frames/s, {byt / us / 1e3:.1f} GB/s algorithmic')
