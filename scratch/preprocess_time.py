"""Host vs device preprocessor on a Food-Kitchen-sized synthetic log (timing only; parity is tests/test_gpu_preprocess.py)."""
import random, time, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from c2dsr_b200 import dataloader as dl
na, nb, L, n = 29207, 34886, 15, 34117
r = np.random.RandomState(0)
seqs = [r.randint(0, na + nb, int(r.randint(4, L + 2))).tolist() for _ in range(n)]
torch.cuda.init(); torch.zeros(1, device="cuda")
out = {}
for name, fn in (("host", lambda: dl.preprocess_train(seqs, na, nb, L, rng=random.Random(1))),
                 ("device", lambda: dl.preprocess_train_device(seqs, na, nb, L, "cuda", rng=random.Random(1)))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter(); x = fn(); torch.cuda.synchronize()
    out["train_" + name + "_s"] = round(time.perf_counter() - t, 4)
ev = seqs[:8000]
for name, fn in (("host", lambda: dl.preprocess_evaluate(ev, na, nb, L, 999, rng=random.Random(1))),
                 ("device", lambda: dl.preprocess_evaluate_device(ev, na, nb, L, 999, "cuda", rng=random.Random(1)))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter(); x = fn(); torch.cuda.synchronize()
    out["eval_" + name + "_s"] = round(time.perf_counter() - t, 4)
out["n_train_seqs"], out["n_eval_seqs"] = n, len(ev)
print(json.dumps(out))
