"""pyflow -- B200-native drop-in for the projection time-step path of
beautifulplanet/PyFlow-Solver (lid-driven cavity).  Same import paths as the
reference for that path: ``pyflow.drivers.simulation_driver.SimulationDriver``,
``pyflow.solvers.solver.step``, ``pyflow.solvers.pressure_solver``,
``pyflow.numerics.core_ops_coherent``, ``pyflow.config.model``,
``pyflow.io.checkpoint``, ``pyflow.cli``.  Compute runs in libpyflow_b200
(hand-written sm_100a CUDA) -- there is no CPU fallback."""
__version__ = "0.1.0"
