"""Loop keyword arguments of the SQD-loop golden cases, shared by make_golden.py and the tests."""

LOOP_CASES = [
    dict(spin_sq=0.0, kw=dict(samples_per_batch=12, num_batches=3, max_iterations=4, symmetrize_spin=True,
                              carryover_threshold=1e-3, seed=7)),
    dict(spin_sq=None, kw=dict(samples_per_batch=15, num_batches=2, max_iterations=3, max_dim=(12, 10),
                               include_configurations=([7, 11], [3]), seed=21)),
    dict(spin_sq=None, kw=dict(samples_per_batch=8, num_batches=4, max_iterations=5, max_dim=9,
                               carryover_threshold=5e-2, seed=3, energy_tol=1e-6, occupancies_tol=1e-3)),
]
