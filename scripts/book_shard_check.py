"""torchrun --nproc-per-node N scripts/book_shard_check.py: a book priced over N ranks (items dealt over
the ranks, one all-reduce) must equal the same book priced by one rank alone, bit for bit."""
import os, sys
sys.path.insert(0, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200"); sys.path.insert(0, ".")
import torch, torch.distributed as dist
from optpricer_b200 import book, _native as nv
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nv.init(local)
rows = book.synthetic_book(300, n_paths=20_000)
many = book.price_book(rows, mc_paths=20_000)                      # >= 4 items per rank: by item
alone = book.price_book(rows, mc_paths=20_000, sharded=False)
few = book.price_book(rows[:5], mc_paths=20_000)                   # by paths
few_alone = book.price_book(rows[:5], mc_paths=20_000, sharded=False)
if dist.get_rank() == 0:
    same = all(a["price"] == b["price"] and a["stderr"] == b["stderr"] for a, b in zip(many, alone))
    close = max(abs(a["price"] - b["price"]) / max(1.0, abs(b["price"])) for a, b in zip(few, few_alone))
    print("world", dist.get_world_size(), "item-sharded book bit-identical to single rank:", same,
          "| path-sharded max rel diff %.2e" % close)
    assert same and close < 1e-5
dist.destroy_process_group()
