"""Drop-in for ``nnmnkwii.baseline`` (GMM-based voice conversion around the MLPG hot path)."""
