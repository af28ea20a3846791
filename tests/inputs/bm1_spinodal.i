# PFHub benchmark 1a (spinodal decomposition): 200^2 periodic Cahn-Hilliard with the double-well
# rho_s (c - c_alpha)^2 (c_beta - c)^2, M = 5, kappa = 2, deterministic cosine initial condition,
# AdamsBashforthMoulton with 1000 substeps per step, time step growing by 1.1.  Same setup as the
# reference's benchmarks/01_spinodal_decomposition/1a_solver.i (BASELINE.json configs[1]); the two stale
# parameters of that file (`history_size`, `spectral_solve_substeps`, no longer declared by the classes)
# and the finite-element Terminator are left out.
[Domain]
  dim = 2
  nx = 200
  ny = 200
  xmax = 200
  ymax = 200
  mesh_mode = DOMAIN
[]

[TensorComputes]
  [Initialize]
    [c]
      type = ParsedCompute
      buffer = c
      extra_symbols = true
      expression = 'c0+epsilon*(cos(0.105*x)*cos(0.11*y)+(cos(0.13*x)*cos(0.087*y))^2+cos(0.025*x-0.15*y)*cos(0.07*x-0.02*y))'
      constant_names = 'c0 epsilon'
      constant_expressions = '0.5 0.01'
    []
    [Mbar]
      type = ReciprocalLaplacianFactor
      buffer = Mbar
      factor = 5
    []
    [kappabarbar]
      type = ReciprocalLaplacianSquareFactor
      buffer = kappabarbar
      factor = -10
    []
  []
  [Solve]
    [mu]
      type = ParsedCompute
      buffer = mu
      expression = 'rho_s*(c-c_alpha)^2*(c_beta-c)^2'
      constant_names = 'rho_s c_alpha c_beta'
      constant_expressions = '5 0.3 0.7'
      derivatives = c
      inputs = c
    []
    [mubar]
      type = ForwardFFT
      buffer = mubar
      input = mu
    []
    [Mbarmubar]
      type = ParsedCompute
      buffer = Mbarmubar
      expression = 'Mbar*mubar'
      inputs = 'Mbar mubar'
    []
    [cbar]
      type = ForwardFFT
      buffer = cbar
      input = c
    []
  []
  [Postprocess]
    [Fgrad]
      type = FFTGradientSquare
      buffer = Fgrad
      input = c
      factor = 1
    []
    [F]
      type = ParsedCompute
      buffer = F
      expression = 'rho_s * (c-c_alpha)^2 * (c_beta-c)^2 + Fgrad'
      constant_names = 'rho_s c_alpha c_beta'
      constant_expressions = '5 0.3 0.7'
      inputs = 'c Fgrad'
    []
  []
[]

[TensorSolver]
  type = AdamsBashforthMoulton
  buffer = c
  reciprocal_buffer = cbar
  linear_reciprocal = kappabarbar
  nonlinear_reciprocal = Mbarmubar
  substeps = 1000
[]

[Postprocessors]
  [min_c]
    type = TensorExtremeValuePostprocessor
    buffer = c
    value_type = MIN
  []
  [max_c]
    type = TensorExtremeValuePostprocessor
    buffer = c
    value_type = MAX
  []
  [F]
    type = TensorIntegralPostprocessor
    buffer = F
  []
  [change]
    type = TensorIntegralChangePostprocessor
    buffer = c
  []
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 1000
  dtmax = 300
  [TimeStepper]
    type = IterationAdaptiveDT
    growth_factor = 1.1
    dt = 1
  []
[]

[Outputs]
  csv = true
  file_base = bm1_spinodal
[]
