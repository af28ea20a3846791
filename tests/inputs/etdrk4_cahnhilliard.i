# Cahn-Hilliard with the ETDRK4Solver: a NON-ZERO nonlinear term, so that all four stage evaluations and the
# phi_1..phi_3 coefficients (including their L*dt == 0 branch at the k = 0 mode: dt, dt^2/2, dt^2/6 as coded in
# src/tensor_solver/ETDRK4Solver.C:84-91) enter the result.  The reference's own ETDRK4 test
# (test/tests/solvers/etdrk4_diffusion.i) has nonlinear_reciprocal = zero and cannot see them.
# kappa and dt are chosen so that |L dt| = O(0.1 .. 1e4) on every k != 0 mode: the reference's phi formulas divide by
# (L dt)^3 and lose all accuracy by cancellation for |L dt| << 1, where no two exp() implementations agree.
# The nonlinear term carries a k-independent part (`+ 0.5*cbar`, a linear growth term) so that its k = 0 mode does not vanish.
[Domain]
  dim = 2
  nx = 32
  ny = 32
  xmax = 4
  ymax = 4
  mesh_mode = DUMMY
[]

[TensorComputes]
  [Initialize]
    [c]
      type = RandomTensor
      buffer = c
      min = 0.44
      max = 0.56
      seed = 0
    []
    [Mbar]
      type = ReciprocalLaplacianFactor
      factor = 0.2
      buffer = Mbar
    []
    [kappabarbar]
      type = ReciprocalLaplacianSquareFactor
      factor = -0.5
      buffer = kappabarbar
    []
  []
  [Solve]
    [mu]
      type = ParsedCompute
      buffer = mu
      expression = '0.1*c^2*(c-1)^2'
      derivatives = c
      inputs = c
    []
    [mubar]
      type = ForwardFFT
      buffer = mubar
      input = mu
    []
    [Nbar]
      type = ParsedCompute
      buffer = Nbar
      expression = 'Mbar*mubar + 0.5*cbar'
      inputs = 'Mbar mubar cbar'
    []
    [cbar]
      type = ForwardFFT
      buffer = cbar
      input = c
    []
  []
[]

[TensorSolver]
  type = ETDRK4Solver
  buffer = c
  reciprocal_buffer = cbar
  linear_reciprocal = kappabarbar
  nonlinear_reciprocal = Nbar
  substeps = 4
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 3
  dt = 0.2
[]
