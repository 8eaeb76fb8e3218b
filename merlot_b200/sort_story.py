"""Zero-shot story unshuffling: the host-side scoring of downstream/sort_story (sort_story half).

`get_zero_shot_logits.py:55-86` runs MerlotModel forward (is_training=False, shuffled_idx_img = 64 + a fixed random order) and
keeps `softmax(allpairs_temporal_logits)[:, 1:]` -- P(same position), P(i before j), P(i after j) for every (sentence i, image j)
pair -- averaged over `duplication_factor` copies.  `score_permutations.py:15-71` then scores each of the n! assignments of
images to sentences and keeps the best.  `temporal_probs` is the first half on top of our MerlotModel mirror; everything else is
plain NumPy and is tested bit for bit against a loop restatement.
"""
from __future__ import annotations

import itertools
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np


def temporal_probs(model, duplication_factor: int = 1, scope_name: str = "lang_viz_temporal"):
    """get_zero_shot_logits.py:68-86 for one modality pair: [batch, n, n, 3] probabilities of (equal, before, after)."""
    import torch
    n, H = model.num_chunks_in_group, model.hidden_size
    h_lang = model.encoder_hidden_states["lang"].reshape(model.B, n, model.lang_chunk_length, H)[:, :, 0]
    h_viz = model.encoder_hidden_states["viz"].reshape(model.B, n, model.viz_chunk_length, H)[:, :, 0]
    xa, xb = (h_lang, h_viz) if scope_name == "lang_viz_temporal" else (h_viz, h_viz)
    logits = model.allpairs_temporal_logits(xa=xa, xb=xb, scope_name=scope_name)
    probs = torch.softmax(logits.float(), -1)[:, 1:]
    return probs.reshape(model.B // duplication_factor, duplication_factor, n, n, 3).mean(1)


def permutation_scores(probs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """score_permutations.py:15-29,57-62 for all permutations at once.  probs [n, n, 3]; sentence i keeps position i, image j is
    assigned position perm[j]; a pair contributes log P(eq) if i == perm[j], log P(before) if i < perm[j], log P(after) otherwise.
    Returns (perms [n!, n] in itertools order, scores [n!])."""
    probs = np.asarray(probs, dtype=np.float64)
    n = probs.shape[0]
    if probs.shape != (n, n, 3):
        raise ValueError(f"probs must be [n, n, 3], got {probs.shape}")
    perms = np.array(list(itertools.permutations(range(n))), dtype=np.int64)  # [n!, n]
    i = np.arange(n)[None, :, None]                                            # sentence positions
    pj = perms[:, None, :]                                                     # image positions under each permutation
    cls = np.where(i == pj, 0, np.where(i < pj, 1, 2))                         # [n!, n, n]
    lp = np.log(probs)
    picked = np.take_along_axis(np.broadcast_to(lp[None], (len(perms), n, n, 3)), cls[..., None], axis=-1)[..., 0]
    return perms, picked.sum((1, 2))


def best_permutation(probs: np.ndarray) -> Tuple[Tuple[int, ...], float]:
    """The permutation the reference keeps (:64): highest score, first in itertools order among ties (stable sort)."""
    perms, scores = permutation_scores(probs)
    k = int(np.argmax(scores))
    return tuple(int(v) for v in perms[k]), float(scores[k])


def spearman_acc(story: Sequence[int]) -> float:  # :32-33 (scipy.stats.spearmanr against the identity; no ties in a permutation)
    s = np.asarray(story, dtype=np.float64)
    n = len(s)
    d = s - np.arange(n)
    return float(1.0 - 6.0 * np.sum(d * d) / (n * (n * n - 1)))


def absolute_distance(story: Sequence[int]) -> float:  # :35-36
    s = np.asarray(story, dtype=np.float64)
    return float(np.mean(np.abs(s - np.arange(len(s)))))


def pairwise_acc(story: Sequence[int]) -> float:  # :39-46
    n = len(story)
    correct = sum(1 for a in range(n) for b in range(a + 1, n) if story[a] < story[b])
    return correct / (n * (n - 1) // 2)


def evaluate(all_probs: Iterable[np.ndarray]) -> Dict[str, float]:
    """score_permutations.py:54-80: best permutation per story, then the three means the script prints."""
    stories: List[Tuple[int, ...]] = [best_permutation(p)[0] for p in all_probs]
    return {"spearman": float(np.mean([spearman_acc(s) for s in stories])),
            "absolute_distance": float(np.mean([absolute_distance(s) for s in stories])),
            "pairwise": float(np.mean([pairwise_acc(s) for s in stories])), "stories": stories}


def write_logits_h5(path: str, predictions: Iterable[Dict], include_images: bool = True) -> int:
    """The writer loop of get_zero_shot_logits.py:105-119: one HDF5 group per story id with `permutation_identity_encode`,
    optional uint8 `images` (255 * image), `sentences`, and `{lang_viz,viz_viz}_probs`; a story id seen twice is skipped like the
    reference's `except ValueError: continue`.  `predictions` yields per-story dicts (what estimator.predict yields there; here
    e.g. {'story_id': id, 'lang_viz_probs': temporal_probs(model)[i].cpu().numpy(), ...}).  Returns the number of groups written.
    Needs h5py (not part of this image: raises ImportError with that message instead of writing another format silently)."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("write_logits_h5 needs h5py (downstream/sort_story/get_zero_shot_logits.py:105 writes HDF5); "
                          "use write_logits_npz for an h5py-free container with the same keys") from e
    n = 0
    with h5py.File(path, "w") as h5:
        for x in predictions:
            key = str(x["story_id"])
            if key in h5:
                continue
            grp = h5.create_group(key)
            grp.create_dataset("permutation_identity_encode", data=np.asarray(x["permutation_identity_encode"]))
            if include_images and "images" in x:
                grp.create_dataset("images", data=(255 * np.asarray(x["images"], dtype=np.float32)).astype(np.uint8))
            grp.create_dataset("sentences", data=np.asarray(x["sentences"]))
            for modality_name in ("lang_viz", "viz_viz"):
                grp.create_dataset(f"{modality_name}_probs", data=np.asarray(x[f"{modality_name}_probs"]))
            n += 1
    return n


def write_logits_npz(path: str, predictions: Iterable[Dict], include_images: bool = False) -> int:
    """Same records as write_logits_h5 in a NumPy .npz (keys `<story_id>/<dataset>`), for environments without h5py;
    read_logits_npz gives back {story_id: {dataset: array}} -- the structure score_permutations.py walks (:77-87)."""
    out, seen = {}, set()
    for x in predictions:
        key = str(x["story_id"])
        if key in seen:
            continue
        seen.add(key)
        out[f"{key}/permutation_identity_encode"] = np.asarray(x["permutation_identity_encode"])
        if include_images and "images" in x:
            out[f"{key}/images"] = (255 * np.asarray(x["images"], dtype=np.float32)).astype(np.uint8)
        out[f"{key}/sentences"] = np.asarray(x["sentences"])
        for modality_name in ("lang_viz", "viz_viz"):
            out[f"{key}/{modality_name}_probs"] = np.asarray(x[f"{modality_name}_probs"])
    np.savez_compressed(path, **out)
    return len(seen)


def read_logits_npz(path: str) -> Dict[str, Dict[str, np.ndarray]]:
    res: Dict[str, Dict[str, np.ndarray]] = {}
    with np.load(path) as z:
        for k in z.files:
            sid, name = k.split("/", 1)
            res.setdefault(sid, {})[name] = z[k]
    return res


def fixed_shuffle_index(batch_size: int, num_chunks: int, seed: int = 1234) -> np.ndarray:
    """The evaluation-time frame order of get_zero_shot_logits.py:55-56: argsort of a FIXED uniform draw per story, + 64 (the
    'hard' index range of temporal_loss, model/modeling.py:635).  The reference draws with tf.random.stateless_uniform(seed=
    [123, 1234]); that generator cannot be reproduced without TensorFlow, so the order differs from the reference's but is
    equally fixed (stated in INTEGRATION.md)."""
    rng = np.random.RandomState(seed)
    return (np.argsort(rng.uniform(size=(batch_size, num_chunks)), 1).astype(np.int32) + 64)
