"""Plain-data descriptions of test fixtures that __graft_entry__.build() pre-compiles kernel plans for.
No imports on purpose: build() must not pull in the test helpers or the oracle."""
U0, Ut, Ux, Uxx = (0, 0), (1, 0), (0, 1), (0, 2)

CASES = [
    # name, d, hidden widths, activation, lhs, rhs, n points, masked
    ("one_hidden_layer", 2, [24], "tanh", [(Ut, 1)], [[(U0, 1)], [(Uxx, 1)], [(U0, 1), (Ux, 1)]], 19, ()),
    ("width_8", 2, [8, 8, 8], "rational", [(Ut, 1)], [[(Ux, 1)], [(Uxx, 1)]], 11, ()),
    ("width_64", 2, [64, 64], "tanh", [(Ut, 1)], [[(Uxx, 1)], [(U0, 2)]], 9, ()),
    ("ragged_50_12_33", 2, [50, 12, 33], "tanh", [(Ut, 1)], [[(Ux, 1)], [(U0, 1), (Uxx, 1)]], 10, ()),
    ("value_only_J1", 2, [16, 16], "tanh", [(U0, 1)], [[(U0, 3)]], 13, ()),
    ("no_rhs_terms_K0", 2, [16, 16], "tanh", [(Ut, 1)], [], 13, ()),
    ("all_terms_masked", 2, [16, 16], "rational", [(Ut, 1)], [[(Ux, 1)], [(Uxx, 1)]], 13, (0, 1)),
    ("high_powers", 2, [16, 16], "tanh", [(Ut, 2)], [[(U0, 5)], [(Ux, 3), (U0, 2)], [(Uxx, 2), (Ux, 1)]], 13, ()),
    ("d3_mixed_xy", 3, [20, 20], "tanh", [((1, 0, 0), 1)], [[((0, 1, 1), 1)], [((0, 2, 0), 1)], [((0, 0, 0), 1), ((0, 0, 1), 1)]], 12, ()),
    ("second_order_time", 2, [16, 16, 16], "rational", [((2, 0), 1)], [[(Uxx, 1)], [((1, 1), 1)]], 12, ()),
]

# (environment knobs, tolerance) of tests/test_gpu_parity.py::test_tensor_pipe_kernel_variants_subprocess
# (default = 14: scaled fp16 forward cross terms + bf16 adjoint cross terms; 6 = the round-1 default with a 3xTF32 forward)
KERNEL_VARIANTS = [({"PDL_CROSS_BF16": "0"}, 1e-5), ({"PDL_CROSS_BF16": "6"}, 1e-5), ({"PDL_ADJ_FUSED": "1"}, 1e-5), ({"PDL_ADJ_FUSED": "0"}, 1e-5),
                   ({"PDL_CROSS_BF16": "7"}, 3e-5)]
