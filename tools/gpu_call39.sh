bash tools/run_gpu_round.sh r05f "k_cov_fill|k_prior_groups|k_predict_fused2|k_assemble_A|k_leaf_q2|k_leaf_cov_fill|k_leaf_gram|k_leaf_ut2|k_fold|k_leaf_chol|k_knot|k_node|k_leaf_linv|k_leaf_qobs"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
