"""Multi-GPU data-parallel parity (pytest -m gpu; skipped on a 1-GPU box): two NCCL ranks, different per-rank batches.
  * every loss of every rank against the WORLD-SIZE-2 oracle restatement (contrastive features gathered over replicas, labels
    shifted by rank*N: model/modeling.py:504-519, utils/model_utils.py:673-707);
  * the gradient the optimizer sees (bucketed, overlapped all-reduce of train.py, 1/world folded into AdamW) against autograd
    of the replica-mean loss (CrossShardOptimizer, utils/optimization.py:241-245) -- including the cross-replica gradient of
    the feature gather (reduce-scatter);
  * bit-identical parameters on both ranks after optimizer steps.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        from merlot_b200.config import NeatConfig
        from merlot_b200.modeling import MerlotModel
        from merlot_b200.optimization import build_optimizer_from_config
        from merlot_b200.params import ParamStore
        from merlot_b200.train import DataParallel
        from oracle import merlot_oracle as O
        from tests.test_gpu_model import synth
        dp = DataParallel("nccl")
        cfg = dict(use_bfloat16=True, hidden_size=128, vocab_size=1000, patch_size=16, spatial_pool_size=2, num_attention_heads=2,
                   num_hidden_layers=2, num_vision_transformer_hidden_layers=4, num_lang_transformer_hidden_layers=2,
                   intermediate_size=256, initializer_range=0.02, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                   max_position_embeddings=64, num_chunks_in_group=2, do_projection=True, do_bias=True, contrastive_size=128,
                   contrast_coef=0.25, contrast_temp=0.05, image_shuffle_prob=0.4, masking_rate=0.2, resnet_layers=[])
        ocfg = dict(type="adam_optimizer", learning_rate=3e-4, num_train_steps=1000, num_warmup_steps=10, weight_decay_rate=0.1,
                    beta_2=0.98, clip_norm=0.0, use_bfloat16_adam=True,
                    param_overrides=[[["LayerNorm", "layer_norm", "GroupNorm", "bias"], {"weight_decay_rate": 0}]])
        params = O.init_params(cfg, seed=1, perturb=0.05)
        params = {k: (v.bfloat16().float() if (k.endswith("kernel") or k.endswith("word_embeddings")) else v) for k, v in params.items()}
        store = ParamStore(cfg, device=dev, optimizer_cfg=ocfg)
        store.load_tf_dict(params)
        batch, nc, Lc = 2, 4, 16
        data = [synth(cfg, batch, nc, Lc, 64, 96, 10 + r) for r in range(world)]  # every rank can rebuild every rank's batch
        B, Lj = batch * nc // 2, Lc * 2
        draws = [O.make_mask_draws(B, Lj, int(Lj * 0.2), cfg["vocab_size"], seed=5 + r) for r in range(world)]
        image, ids, shuf, vid = data[rank]
        m = MerlotModel(cfg, is_training=False, use_tpu=False, image=image.to(dev), input_ids=ids.to(dev), mask_input=True,
                        shuffled_idx_img=shuf.to(dev), params=store, mask_draws=draws[rank], save_for_backward=True, dist=dp)
        ll, _ = m.mask_loss()
        cl, cinfo = m.contrastive_loss()
        tl, _ = m.temporal_loss(shuf.to(dev), vid.to(dev))
        my = torch.tensor([float(ll), float(cl), float(tl)], device=dev)
        mine_masks = torch.cat([m.lang_mask_info["masked_ids"].reshape(-1), m.lang_mask_info["masked_idx"].reshape(-1)]).int()
        all_masks = [torch.empty_like(mine_masks) for _ in range(world)]
        dp.dist.all_gather(all_masks, mine_masks)
        # the data-parallel backward exactly as train.py's train_op runs it (bucketed all-reduce, no optimizer yet)
        store.g.zero_()
        groups, vit_ranges = store.vit_buckets(2)
        pending, vit_pending = [], []
        m.backward(on_non_vit_grads_ready=lambda: pending.extend(dp.all_reduce_ranges_async(store.g, store.rest_ranges)),
                   vit_layer_groups=groups,
                   on_vit_group_done=lambda k: vit_pending.append(dp.all_reduce_ranges_async(store.g, vit_ranges[k])) if k + 1 < len(groups) else None)
        vit_pending.append(dp.all_reduce_ranges_async(store.g, vit_ranges[-1]))
        dp.wait_all(pending)
        for hs in vit_pending:
            dp.wait_all(hs)
        torch.cuda.synchronize()
        g_mean = {k: v / world for k, v in store.to_tf_dict("g").items()}
        out = {"rank": rank, "losses": my.cpu().tolist()}
        if rank == 0:  # world-size-2 oracle on one autograd graph
            leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
            oms = []
            for r in range(world):
                im, idr, sh, _ = data[r]
                mk = all_masks[r].cpu()
                gm = {"masked_ids": mk[:B * Lj].reshape(B, Lj), "masked_idx": mk[B * Lj:].reshape(B, -1)}
                oms.append(O.MerlotOracle(cfg, leaf, im, idr, mask_input=True, shuffled_idx_img=sh, mask_override=gm))
            mean_loss, per = O.pretrain_losses_replicas(oms, [d[2] for d in data], [d[3] for d in data])
            mean_loss.backward()
            worst = {}
            for k, v in leaf.items():
                if v.grad is None or float(v.grad.norm()) < 1e-7:
                    continue
                worst[k] = rel(g_mean[k], v.grad)
            out["oracle_per_rank_total"] = [float(x) for x in per]
            out["grad_worst"] = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
            out["grad_max"] = max(worst.values())
            _, c0 = O.contrastive_loss_replicas(oms, 0)
            out["oracle_contr0"] = float(c0["loss_all"])
        # optimizer steps through the real train_op path: parameters must stay bit-identical across ranks
        store.g.zero_()
        from merlot_b200.train import model_fn_builder
        ncfg = NeatConfig.from_dict({"data": {"num_chunks": 4, "chunk_text_len": Lc}, "model": dict(cfg, image_size=[64, 96], hidden_dropout_prob=0.1),
                                     "optimizer": ocfg, "device": {"use_tpu": False, "output_dir": "/tmp/x"}})
        fn = model_fn_builder(ncfg, store=store, dist=dp, device=dev, vit_grad_buckets=3)
        feats = {"images": image.to(dev).bfloat16(), "input_ids": ids.to(dev), "shuffled_idx_img": shuf.to(dev), "video_src_ids": vid.to(dev)}
        for _ in range(3):
            fn(feats).train_op()
        torch.cuda.synchronize()
        pl = [torch.empty_like(store.p) for _ in range(world)]
        dp.dist.all_gather(pl, store.p)
        out["params_identical"] = all(torch.equal(pl[0], x) for x in pl[1:])
        out["step"] = store.global_step
        q.put(out)
        dp.barrier()
    except Exception as e:  # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put({"rank": rank, "error": traceback.format_exc()})
        raise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_losses_gradients_and_parameters():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in ps:
        p.join(timeout=120)
    for r in res:
        assert "error" not in r, r.get("error")
    res = {r["rank"]: r for r in res}
    r0 = res[0]
    for r in range(world):
        total = sum(res[r]["losses"])
        ref = r0["oracle_per_rank_total"][r]
        assert abs(total - ref) <= 1e-3 * abs(ref), (r, total, ref)          # every rank's loss vs the 2-replica oracle
    assert abs(res[0]["losses"][1] - r0["oracle_contr0"]) <= 2e-3 * abs(r0["oracle_contr0"])  # the gathered contrastive term itself
    assert r0["grad_max"] < 4e-2, r0["grad_worst"]                             # same bar as the single-GPU model test
    assert all(res[r]["params_identical"] for r in range(world)) and all(res[r]["step"] == 3 for r in range(world))
