from torch.utils.data import DataLoader  # noqa: F401  (PyG's DataLoader collates ints to an int64 tensor, as torch's does)
