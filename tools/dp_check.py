"""Data-parallel parity check (run with torchrun on N GPUs): N ranks each take 1/N of one global batch; the
result (loss, every gradient group, updated parameters) must match the float64 oracle run single-device on the
whole batch - which needs the global 1/B loss scale, the cross-rank batch-norm statistics and clip-after-allreduce."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist
from oracle import Config, GROUPS, Oracle, init_params, init_bn_stats, make_batch
from parity_util import make_handle, rel_err
from aocr import dist as adist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
GB = 4 * world
cfg = Config(batch_size=GB // world, max_encoder_l=30, max_decoder_l=12)
full = make_batch(GB, 100, 7, seed=17)
params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
from aocr.capi import AocrConfig, Handle
c = AocrConfig(batch_size=cfg.batch_size, max_encoder_l=30, max_decoder_l=12, encoder_num_hidden=512, encoder_num_layers=1,
               decoder_num_layers=2, target_vocab_size=39, target_embedding_size=20, input_feed=1, dropout=0.0,
               learning_rate=0.1, dp_rank=rank, dp_world=world, global_batch=GB, gemm_mode=int(os.environ.get("AOCR_GEMM_MODE", "0")))
h = Handle(c, local)
for i, g in enumerate(GROUPS):
    h.set_params(i, params[g])
for i, k in enumerate(("bn3", "bn5", "bn7")):
    h.set_bn_stats(i, *bn[k])
gs = adist.attach(h, local) if os.environ.get("AOCR_DP_HOOK") else adist.attach_native(h, local)
sl = slice(rank * GB // world, (rank + 1) * GB // world)
loss_local = h.forward_backward(full["images"][sl], full["targets"][sl], full["targets_eval"][sl])
t = torch.tensor([loss_local], dtype=torch.float64, device="cuda")
dist.all_reduce(t)
grads = [h.get_grads(i) for i in range(5)]
h.sgd_update(0.1, 5.0)
newp = [h.get_params(i) for i in range(5)]
kinds = [k for k, _ in gs.log.calls] if hasattr(gs, "log") else "native NCCL"
if rank == 0:
    ocfg = Config(batch_size=GB, max_encoder_l=30, max_decoder_l=12)
    orc = Oracle(ocfg, params, bn)
    lo, go, _ = orc.forward_backward(full["images"], full["targets"], full["targets_eval"])
    orc.sgd_update(go, 0.1)
    po = orc.flat_params()
    print(f"world {world}: loss dp={float(t[0]):.6f} oracle={lo:.6f} rel={abs(float(t[0]) - lo) / lo:.2e}; hook calls {kinds}")
    ok = abs(float(t[0]) - lo) / lo < 1e-3
    for i, g in enumerate(GROUPS):
        e_g = np.linalg.norm(grads[i].astype(np.float64) - go[g]) / np.linalg.norm(go[g])
        e_p = rel_err(newp[i], po[g])
        print(f"  {g:8s} grad L2-rel {e_g:.2e}   updated-param rel {e_p:.2e}")
        ok = ok and e_g < (2e-2 if g == "cnn" else 2e-3) and e_p < 1e-3
    print("DP_PARITY_OK" if ok else "DP_PARITY_FAIL")
h.close()
dist.destroy_process_group()
