"""Second, independent CPU restatement of the token-mask sampler - TEST INFRASTRUCTURE, NOT PRODUCT (same rules as
oracle/multimae_oracle.py: only tests/ may import it).

Plain numpy integer / float32 arithmetic, no torch: MultiMAE.generate_random_masks (multimae/multimae.py:189-216) as a
pure function of its random draws.  It exists so that the bit-exact index parity of the CUDA sampler is not anchored on
torch's argsort alone: the torch oracle (multimae_oracle.sample_masks), this file and the fixtures recorded from the live
reference (tests/golden/sampler_*.pt) must agree index for index."""
import numpy as np


def sample_masks(shares, noises, noise_all, num_encoded):
    """shares [B, T] float32 Dirichlet draw; noises[t] [B, N_t] and noise_all [B, sum N_t] float32 uniform draws.
    Returns (task_masks list of [B, N_t] int64 with 0 = visible, ids_keep [B, num_encoded] int64, ids_restore [B, sum N_t])."""
    shares = np.asarray(shares, dtype=np.float32)
    # :189  samples_per_task = (task_sampling_dist * num_encoded_tokens).round().long()  - float32 product, round half to even
    per_task = np.rint(shares * np.float32(num_encoded)).astype(np.int64)
    masks = []
    for t, noise in enumerate(noises):
        noise = np.asarray(noise, dtype=np.float32)
        order = np.argsort(noise, axis=1, kind="stable")                   # :196  ids_arange_shuffle
        # :197-200  mask = where(arange < k_t, 0, 1) gathered at ids_arange_shuffle: position j is visible iff the index
        # of the j-th smallest noise value is below k_t
        masks.append((order >= per_task[:, t:t + 1]).astype(np.int64))
    mask_all = np.concatenate(masks, axis=1)
    key = mask_all.astype(np.float32) + np.asarray(noise_all, dtype=np.float32)      # :204  float32 sum, as torch computes it
    ids_shuffle = np.argsort(key, axis=1, kind="stable")
    ids_restore = np.argsort(ids_shuffle, axis=1, kind="stable")           # :205
    ids_keep = ids_shuffle[:, :num_encoded]                                # :206
    final = np.ones_like(mask_all)                                         # :209-212  exactly num_encoded zeros per row
    final[:, :num_encoded] = 0
    final = np.take_along_axis(final, ids_restore, axis=1)
    splits = np.cumsum([np.asarray(n).shape[1] for n in noises])[:-1]
    return list(np.split(final, splits, axis=1)), ids_keep.astype(np.int64), ids_restore.astype(np.int64)
