"""Console + file logger with the method names ``main.py`` / ``Trainer`` call on the reference's
``utils.noter.Noter`` (log_msg, log_train, log_evaluate, log_final_result).  Logging is outside
the hot path; the layout of the lines is this package's own."""
from __future__ import annotations

import os
import time

HEAD = ("Improve", "hr5_a", "hr20_a", "mrr5_a", "mrr20_a", "ndcg5_a", "ndcg20_a",
        "hr5_b", "hr20_b", "mrr5_b", "mrr20_b", "ndcg5_b", "ndcg20_b")


class Noter(object):
    def __init__(self, args, quiet: bool = False):
        self.args, self.quiet = args, quiet
        self.f_log = None
        path_log = getattr(args, "path_log", None)
        if path_log and getattr(args, "rank", 0) == 0:
            os.makedirs(path_log, exist_ok=True)
            stamp = time.strftime("%m-%d-%H%M%S", time.localtime())
            self.f_log = os.path.join(path_log, f"{getattr(args, 'data', 'run')}-{stamp}-{args.n_gnn}-{args.n_attn}-"
                                                f"{args.n_head}-{args.lr}-{args.l2}.txt")
        self.log_msg(f"[c2dsr_b200] dataset={getattr(args, 'dataset', '?')} d={args.d_latent} n_gnn={args.n_gnn} "
                     f"n_attn={args.n_attn} n_head={args.n_head} lr={args.lr:.2e} l2={args.l2:.2e}")

    def log_msg(self, msg):
        if getattr(self.args, "rank", 0) != 0:
            return
        if not self.quiet:
            print(msg)
        if self.f_log:
            with open(self.f_log, "a") as out:
                print(msg, file=out)

    def log_train(self, loss_tr, loss_rec, loss_mi, t_gap):
        self.log_msg(f"\t| train | loss {loss_tr:.4f} | rec {loss_rec:.4f} | mi {loss_mi:.4f} | time {t_gap:.1f}s |")

    def log_evaluate(self, mode, res):
        self.log_msg(f"\t| {mode:5} | " + " | ".join(f"{h} {v:+.4f}" if i == 0 else f"{h} {v:.4f}"
                                                      for i, (h, v) in enumerate(zip(HEAD, res))) + " |")

    def log_final_result(self, epoch, imp_val_best, res):
        self.log_msg(f"\n[done] stopped at epoch {epoch}; best valid improvement {imp_val_best:+.4f}")
        self.log_evaluate("test", res)
