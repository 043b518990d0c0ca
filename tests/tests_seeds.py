# seeds used by tests/golden/make_golden.py (CASES[...]["seed"])
SEEDS = {"c1_mlp_discrete": 1, "c3_gru_multidiscrete": 2, "c4_gru_smac": 3, "c5_mlp_switches": 4, "naive_rnn_ptl": 5,
         "c5_h512_hanabi": 6, "c2_mlp_n128": 7}
