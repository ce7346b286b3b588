"""collate_fn -- reference: ppvector/data_utils/collate_fn.py:5-23: zero-pad the [T,F] features of a batch to the
longest, return (features [B,Tmax,F], labels [B] int64, input_lens [B] int64)."""
import torch


def collate_fn(batch):
    max_len = max(sample[0].shape[0] for sample in batch)
    feat_dim = batch[0][0].shape[1]
    device = batch[0][0].device
    features = torch.zeros((len(batch), max_len, feat_dim), dtype=torch.float32, device=device)
    labels, input_lens = [], []
    for i, (tensor, label) in enumerate(batch):
        n = tensor.shape[0]
        features[i, :n, :] = tensor
        labels.append(int(label))
        input_lens.append(n)
    return features, torch.tensor(labels, dtype=torch.int64), torch.tensor(input_lens, dtype=torch.int64)
