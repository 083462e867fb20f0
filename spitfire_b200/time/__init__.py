"""Host control logic of the implicit time integration used on the hot path (mirror of spitfire.time):
`odesolve` (integrator.py), `KennedyCarpenterS6P4Q3` (methods.py), `SimpleNewtonSolver` (nonlinear.py),
`PIController` / `ConstantTimeStep` (stepcontrol.py). Only what HomogeneousReactor and Flamelet use is provided;
the reference's other explicit/implicit steppers are a generic ODE library outside the chemistry path."""
from spitfire_b200.time.integrator import odesolve, FailedODESolveException  # noqa: F401
from spitfire_b200.time.methods import KennedyCarpenterS6P4Q3, BackwardEulerS1P1Q1, StepOutput  # noqa: F401
from spitfire_b200.time.nonlinear import SimpleNewtonSolver, SolverOutput, finite_difference_jacobian  # noqa: F401
from spitfire_b200.time.stepcontrol import PIController, ConstantTimeStep  # noqa: F401
