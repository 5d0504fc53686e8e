"""CPU restatement of the reference's legacy pose-feature matcher (TEST INFRASTRUCTURE).

Follows codebook/Speech2GestureMatching/GestureKNN.py:70-284, the `GestureKNN` class: `init_frame` (:89-97),
`search_motion` (:100-152: per 8-frame step, the nearest acceptable frame of EVERY database sequence by L2 distance of
the 96-d pose feature, then rank(pose distance) + rank(audio cosine distance), take the `desired_k`-th candidate),
`search_pose_cands` (:155-214: ascending walk over the argsort of the frame distances, skipping exact zeros, frames
closer than `step_sz` to the end and frames whose control mask is not set at both ends; the LAST element of the
sorted order is never examined, :177), `search_fake_motion` / `search_fake_pose_cands` (:217-284: the same walk on
the audio cosine distance alone, no feedback).  The arithmetic calls are the reference's own (np.linalg.norm per
frame, sklearn paired cosine, np.argsort), so on one machine the results are identical to the reference class.

Pinned against the reference class imported in place (tests/test_legacy_cpu.py) and tests/golden/legacy_*.npz.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
from sklearn.metrics.pairwise import paired_distances


def frame_candidate(dist, mask_row, step_sz):
    """The walk of GestureKNN.py:173-199 over one sequence's frame distances -> (frame, distance) or None."""
    n = len(dist)
    order = np.argsort(dist)
    for f in order[:n - 1]:                                  # `while ctr < len(sorted) - 1` (:177)
        d = dist[f]
        if d == 0.:                                          # :184
            continue
        if f > n - step_sz:                                  # :188
            continue
        if mask_row[f] + mask_row[f + step_sz - 1] != 2:     # :192
            continue
        return int(f), float(d)
    return None


def pose_candidates(feat_train, mask, query, lo, hi, metric, step_sz):
    """For every database sequence with a non-empty mask: nearest acceptable frame.
    Returns (sequence ids, frames, distances) of the sequences that have one, in sequence order."""
    seqs, frames, dists = [], [], []
    for k in range(feat_train.shape[0]):
        if mask[k].sum() == 0:                               # :161
            continue
        rows = feat_train[k, :, lo:hi]
        if metric == "l2":
            dist = [np.linalg.norm(query - rows[l]) for l in range(rows.shape[0])]            # :169-171
        else:
            dist = [paired_distances([query], [rows[l]], metric="cosine")[0] for l in range(rows.shape[0])]   # :258
        got = frame_candidate(np.array(dist), mask[k], step_sz)
        if got is not None:
            seqs.append(k)
            frames.append(got[0])
            dists.append(got[1])
    return np.array(seqs, dtype=np.int64), np.array(frames, dtype=np.int64), np.array(dists)


def search_motion(feat_train, motn_train, mask, feat_test, desired_k, init, n_aud=112, n_body=96, step_sz=8):
    """GestureKNN.search_motion (:100-152) with the initial (sequence, frame) given.  feat_test [n_aud, n_frames].
    Returns (pred_motion [n_joints, n_frames], chosen [(sequence, frame)] per step)."""
    n_frames = feat_test.shape[-1]
    ft = np.concatenate((feat_test[:, 0:1], feat_test), axis=1)
    ft = np.concatenate((ft, np.zeros((n_body, ft.shape[1]))), axis=0)
    ft[n_aud:, 0] = feat_train[init[0], init[1], n_aud:]
    pred = np.zeros((motn_train.shape[2], n_frames + 1))
    chosen = []
    j = 1
    while j < n_frames:
        seqs, frames, pd = pose_candidates(feat_train, mask, ft[n_aud:, j - 1], n_aud, n_aud + n_body, "l2", step_sz)
        ad = np.array([paired_distances([ft[:n_aud, j]], [feat_train[k, f, :n_aud]], metric="cosine")[0]
                       for k, f in zip(seqs, frames)])       # :127-132
        combined = pd.argsort().argsort() + ad.argsort().argsort()        # :135-138
        pick = np.argsort(combined)[desired_k]               # :139, :144
        k, f = int(seqs[pick]), int(frames[pick])
        ft[n_aud:, j:j + step_sz] = feat_train[k, f:f + step_sz, n_aud:].T
        pred[:, j:j + step_sz] = motn_train[k, f:f + step_sz, :].T
        chosen.append((k, f))
        j += step_sz
    return pred[:, 1:], chosen


def search_fake_motion(feat_train, motn_train, mask, feat_test, desired_k, n_aud=112, step_sz=8):
    """GestureKNN.search_fake_motion (:217-243)."""
    n_frames = feat_test.shape[-1]
    pred = np.zeros((motn_train.shape[2], n_frames))
    chosen = []
    j = 0
    while j < n_frames:
        seqs, frames, pd = pose_candidates(feat_train, mask, feat_test[:n_aud, j], 0, n_aud, "cosine", step_sz)
        pick = np.argsort(pd.argsort().argsort())[desired_k]              # :230-232
        k, f = int(seqs[pick]), int(frames[pick])
        pred[:, j:j + step_sz] = motn_train[k, f:f + step_sz, :].T
        chosen.append((k, f))
        j += step_sz
    return pred, chosen


def picks_are_unambiguous(feat_train, mask, feat_test, desired_k, init, fake, n_aud=112, n_body=96, step_sz=8):
    """True when, at every step of the run, the candidate at position desired_k of the rank-sum order holds a sum that
    no other candidate shares and no two candidates have equal distances: then the result does not depend on how an
    unstable sort orders equal keys.  Used by tests/golden/make_golden_legacy.py to keep platform-defined tie
    orders out of the golden vectors."""
    n_frames = feat_test.shape[-1]
    if fake:
        for j in range(0, n_frames, step_sz):
            seqs, frames, pd = pose_candidates(feat_train, mask, feat_test[:n_aud, j], 0, n_aud, "cosine", step_sz)
            if len(np.unique(pd)) != len(pd):
                return False
        return True
    ft = np.concatenate((feat_test[:, 0:1], feat_test), axis=1)
    pose = feat_train[init[0], init[1], n_aud:n_aud + n_body]
    for j in range(1, n_frames, step_sz):
        seqs, frames, pd = pose_candidates(feat_train, mask, pose, n_aud, n_aud + n_body, "l2", step_sz)
        ad = np.array([paired_distances([ft[:n_aud, j]], [feat_train[k, f, :n_aud]], metric="cosine")[0]
                       for k, f in zip(seqs, frames)])
        if len(np.unique(pd)) != len(pd) or len(np.unique(ad)) != len(ad):
            return False
        combined = pd.argsort().argsort() + ad.argsort().argsort()
        pick = np.argsort(combined, kind="stable")[desired_k]
        if (combined == combined[pick]).sum() != 1:
            return False
        pose = feat_train[seqs[pick], frames[pick] + step_sz - 1, n_aud:n_aud + n_body]
    return True
