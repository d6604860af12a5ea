"""Module names of the reference's ``VLAAttacker/white_patch`` package (UADA, UPA, TMA, UADA_ddp, appply_random_transform)."""
