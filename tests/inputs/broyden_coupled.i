# Two coupled Allen-Cahn-type fields integrated with the implicit BroydenSolver (per-wavevector 2x2 inverse
# Jacobian estimate).  The reference ships no gold file for this solver (it is used by
# benchmarks/02_oswald_ripening/2a_broyden.i); the host test compares with the oracle's restatement of
# src/tensor_solver/BroydenSolver.C on the same problem.
[Domain]
  dim = 2
  nx = 32
  ny = 32
  xmax = '${fparse 2*pi}'
  ymax = '${fparse 2*pi}'
  mesh_mode = DUMMY
[]

[TensorComputes]
  [Initialize]
    [u]
      type = ParsedCompute
      buffer = u
      expression = '0.5+0.1*sin(x)*sin(y)'
      extra_symbols = true
      expand = REAL
    []
    [v]
      type = ParsedCompute
      buffer = v
      expression = '0.4+0.1*cos(x)*cos(2*y)'
      extra_symbols = true
      expand = REAL
    []
    [Lu]
      type = ReciprocalLaplacianFactor
      buffer = Lu
      factor = 0.1
    []
    [Lv]
      type = ReciprocalLaplacianFactor
      buffer = Lv
      factor = 0.05
    []
  []
  [Solve]
    [ub]
      type = ForwardFFT
      buffer = ub
      input = u
    []
    [vb]
      type = ForwardFFT
      buffer = vb
      input = v
    []
    [fu]
      type = ParsedCompute
      buffer = fu
      expression = '-(u^3-u) - 0.3*v'
      inputs = 'u v'
    []
    [fub]
      type = ForwardFFT
      buffer = fub
      input = fu
    []
    [fv]
      type = ParsedCompute
      buffer = fv
      expression = '-(v^3-v) - 0.3*u'
      inputs = 'u v'
    []
    [fvb]
      type = ForwardFFT
      buffer = fvb
      input = fv
    []
  []
[]

[TensorSolver]
  type = BroydenSolver
  buffer = 'u v'
  reciprocal_buffer = 'ub vb'
  linear_reciprocal = 'Lu Lv'
  nonlinear_reciprocal = 'fub fvb'
  substeps = 2
  # a fixed number of Broyden iterations (tolerances 0): both implementations follow the same path, so the
  # comparison is at round-off level instead of at the level of the stopping tolerance
  max_iterations = 12
  relative_tolerance = 0
  absolute_tolerance = 0
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 2
  dt = 0.05
[]
