"""a7: the per-epoch shuffle of the reference's training loop (DataLoader(shuffle=True), compare_models.py:291-299) is
replayed bit for bit by HPF_PyTorch.fit_epochs -- with the permutations prefetched by host threads or drawn epoch by
epoch -- and leaves the global generator where the DataLoader leaves it."""
import torch

from prob_matrix_factorization_b200.hpf_pytorch import _EpochPermutations


class _Rows(torch.utils.data.Dataset):          # the scripts' SimpleDataset: index -> (u, i, rating)
    def __init__(self, n):
        self.idx = torch.arange(n)

    def __len__(self):
        return len(self.idx)

    def __getitem__(self, k):
        return self.idx[k]


def test_epoch_permutations_equal_the_dataloader_order():
    n, epochs, batch = 2500, 3, 512
    torch.manual_seed(1234)
    loader = torch.utils.data.DataLoader(_Rows(n), batch_size=batch, shuffle=True)
    want = [torch.cat([b for b in loader]) for _ in range(epochs)]
    after = torch.rand(3)
    for prefetch in (True, False):
        torch.manual_seed(1234)
        got = list(_EpochPermutations(n, epochs, "cpu", prefetch=prefetch))
        assert len(got) == epochs and all(torch.equal(a, b) for a, b in zip(want, got)), prefetch
        assert torch.equal(after, torch.rand(3))          # the same number of draws was taken from the global generator
