"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product package svae_b200).

CPU restatements of the reference's (mattjj/svae) local E-step algorithms, used solely as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

  lds_numpy.py   fp64 NumPy restatement of svae/lds/gaussian.py + svae/lds/lds_inference.py (batched)
  lds_torch.py   the same algorithm in torch.float64 on CPU, so autograd supplies the VJP oracle
  hmm_numpy.py   fp64 NumPy restatement of svae/hmm (log-space forward, E-step, Viterbi)
  gmm_numpy.py   fp64 NumPy restatement of svae/models/gmm.py local_meanfield (+ the distributions it uses)
  slds_numpy.py  fp64 NumPy restatement of svae/models/slds_svae.py optimize_local_meanfield (composed of the above)
  ref.py         loader for oracle/_ref: the reference's own Cython kernels compiled by build_ref.py
  build_ref.py   the build recipe for oracle/_ref

Parity status: PINNED against outputs of the reference itself (oracle/_ref, compiled here from
/root/reference) -- see tests/test_oracle_vs_reference.py and tests/golden/.  The reference's own
tests hold no golden vectors for this path.
"""
