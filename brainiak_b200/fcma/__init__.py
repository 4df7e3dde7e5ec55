"""Full correlation matrix analysis (FCMA) on B200 — the correlation hot path of
``brainiak.fcma`` (VoxelSelector / Classifier / compute_correlation) behind the reference's API."""
