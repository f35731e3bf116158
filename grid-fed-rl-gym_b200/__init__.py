"""B200-native batched GridEnvironment.step. Import as ``grid_fed_rl_b200``."""
