"""Sharding of the path across the GPUs of one box: games / batches are independent units,
rank r owns engine r and its own games, there is NO data-path collective -- `torch.distributed` is only the plumbing
for the start barrier and for adding up the per-shard counters (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Dict, Optional


class Shards:
    def __init__(self, backend: Optional[str] = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.device = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            backend = backend or "nccl"
            if backend == "nccl":
                # with NCCL_DEBUG=VERSION (this image's default) NCCL printf()s a banner on STDOUT, but the benchmarks'
                # stdout is ONE JSON line: say the version on stderr instead; any other NCCL_DEBUG setting is left alone
                if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                    os.environ["NCCL_DEBUG"] = "NONE"
                    if self.rank == 0:
                        import sys
                        print("NCCL version " + ".".join(str(v) for v in torch.cuda.nccl.version()), file=sys.stderr)
                torch.cuda.set_device(self.local_rank)
                self.device = torch.device("cuda", self.local_rank)
                dist.init_process_group("nccl", device_id=self.device)
            else:
                self.device = torch.device("cpu")
                dist.init_process_group(backend)
            self.dist = dist

    def seed(self, base: int) -> int:
        """Every shard plays its own games."""
        return base + self.rank

    def barrier(self) -> None:
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x: float, op: str) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def sum(self, x: float) -> float:
        return self._reduce(x, "SUM")

    def max(self, x: float) -> float:
        return self._reduce(x, "MAX")

    def selfplay_totals(self, stats: Dict[str, float]) -> Dict[str, float]:
        """Whole-job self-play figures: counters add up, the clock is the slowest shard's."""
        seconds = self.max(stats["seconds"])
        moves, evals, rounds = self.sum(stats["moves"]), self.sum(stats["evals"]), self.sum(stats["rounds"])
        return {"moves": moves, "evals": evals, "rounds": rounds, "seconds": seconds,
                "games_finished": self.sum(stats["games_finished"]), "predictor_seconds": self.sum(stats["eval_seconds"]),
                "moves_per_s": moves / seconds, "nn_evals_per_s": evals / seconds, "mean_device_batch": evals / max(rounds, 1.0)}

    def close(self) -> None:
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None
