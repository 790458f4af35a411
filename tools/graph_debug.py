import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_graph import Net
from videotransformer_pytorch_b200.graph import GraphedTrainStep

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

for train_mode in (False, True):
    torch.manual_seed(0)
    net = Net().cuda()
    net.train(train_mode)
    x = torch.randn(2, 4, 3, 48, 48).cuda(); y = torch.tensor([1, 7]).cuda()
    x2 = torch.randn(2, 4, 3, 48, 48).cuda()
    def eager(seed):
        for p in net.parameters(): p.grad = None
        torch.manual_seed(seed)
        l = net(x2, y); l.backward()
        return float(l.detach()), {n: p.grad.clone() for n, p in net.named_parameters()}
    l1, g1 = eager(5)
    l2, g2 = eager(5)
    print(f'train={train_mode} eager determinism: loss {l1} {l2} worst', max(rel(g1[n], g2[n]) for n in g1))
    step = GraphedTrainStep(net, (x, y))
    for rep in range(2):
        torch.manual_seed(5)
        lg = float(step(x2, y).detach())
        torch.cuda.synchronize()
        gg = {n: p.grad.clone() for n, p in net.named_parameters()}
        bad = [(n, rel(gg[n], g1[n])) for n in g1 if rel(gg[n], g1[n]) > 1e-4]
        print(f'train={train_mode} graph rep{rep}: loss {lg} vs {l1}; {len(bad)}/{len(g1)} params mismatch')
        for n, e in bad[:60]:
            print(f'    {n}: {e:.3e}')
        l3, g3 = eager(5)
        bad = [(n, rel(g3[n], g1[n])) for n in g1 if rel(g3[n], g1[n]) > 1e-4]
        print(f'train={train_mode} eager AFTER graph rep{rep}: loss {l3}; {len(bad)}/{len(g1)} params mismatch', bad[:8])
