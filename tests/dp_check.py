"""Data-parallel parity: N ranks on the global batch == one rank on the same batch (dropout 0).
Run once with python (writes /tmp/dp_ref.pt), then under torchrun (compares)."""
import sys, os, argparse
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import Golden, rel_err
from c2dsr_b200 import dist as cdist
from c2dsr_b200.dataloader import BatchLoader, CDSRDataset
from c2dsr_b200.trainer import Trainer

class Quiet:
    def log_train(self, *a): pass
    def log_msg(self, *a): pass

rank, world, local_rank = cdist.init_from_env("nccl")
torch.cuda.set_device(local_rank)
g = Golden("tiny_default")
hp = dict(g.hp)
args = argparse.Namespace(**hp)
args.device = torch.device("cuda", local_rank)
# two passes over 71 samples (batches 32, 32, 7: the last one splits unevenly, 4 + 3 on two ranks), then one over 65
# (32, 32, 1: a remainder shorter than the world, which every rank processes at weight 1 / world)
def make_loader(n):
    ds = CDSRDataset.from_fields([g.z["train_fields"][:n, i] for i in range(14)], "train", hp["len_max"])
    return BatchLoader(ds, hp["batch_size"], rank=rank, world_size=world, len_rec=hp["len_rec"],
                       ignore=(hp["n_item_a"], hp["n_item_b"]))
loader = make_loader(71)
torch.manual_seed(hp["seed"])
tr = Trainer.from_parts(args, Quiet(), (loader, None, None), g.adj("share"), g.adj("spec"))
tr.model.load_state_dict({k: v.cuda() for k, v in g.group("init").items()})
tr.model.train(); tr.optimizer.zero_grad()
losses = []
for epoch, ld in enumerate((loader, loader, make_loader(65))):
    for batch in ld:
        losses.append([float(x) for x in tr.train_step(batch)])
state = {k: v.detach().cpu() for k, v in tr.model.state_dict().items()}
if world == 1:
    torch.save({"losses": losses, "state": state}, "/tmp/dp_ref.pt")
    print("reference saved:", len(losses), "steps", losses[0], losses[-1], "graphs", bool(tr._graphs))
else:
    ref = torch.load("/tmp/dp_ref.pt")
    worst = max(abs(a - b) / abs(b) for la, lb in zip(losses, ref["losses"]) for a, b in zip(la, lb))
    d = hp["d_latent"]
    # q / k rows of in_proj: gradients are rounding noise (softmax over a single allowed key, Q18) -> compare v rows
    sel = lambda k, t: t[2 * d:] if "in_proj" in k else t
    wk = max((rel_err(sel(k, state[k]), sel(k, ref["state"][k])), k) for k in state if not k.endswith("attn_mask"))
    if rank == 0:
        print(f"dp{world}: {len(losses)} steps, worst loss rel diff {worst:.2e}, worst weight rel err {wk[0]:.2e} ({wk[1]}), graphs {bool(tr._graphs)}")
    assert len(losses) == len(ref["losses"]) and worst < 1e-4 and wk[0] < 1e-3, (worst, wk)
    torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
