"""TEST INFRASTRUCTURE ONLY -- minimal restatement of the pieces of `rl-games==1.1.4`
(/root/reference/requirements.txt:4, not vendored, not installable offline) that the reference's
ase/learning/*.py import.  Written from recollection of the public rl_games 1.1.4 sources
(SURVEY.md Appendix A.1); arithmetic-bearing pieces: RunningMeanStd, ModelA2CContinuousLogStd,
A2CBuilder MLP/initialisers, torch_ext.{policy_kl,normalization_with_masks,mean_list},
swap_and_flatten01, ExperienceBuffer, PPODataset, DefaultRewardsShaper.
Pinned against the shipped checkpoints' identities (tests/test_oracle_cpu.py)."""
