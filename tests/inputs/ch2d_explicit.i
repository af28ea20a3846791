# 50x50 Cahn-Hilliard spinodal decomposition integrated EXPLICITLY: one ParsedCompute assembles the
# reciprocal-space time derivative Mbar*mubar - Mkappabarbar*cbar, ForwardEulerSolver advances it with
# 50 substeps per step.  Same setup as the reference's test/tests/cahnhilliard/cahnhilliard_explicit.i
# (gold cahnhilliard_explicit_out.e).  ch2d_explicit_smooth.i adds the de-aliasing filter.
[Domain]
  dim = 2
  nx = 50
  ny = 50
  xmax = 3
  ymax = 3
  mesh_mode = DOMAIN
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
    [mu_init]
      type = ConstantTensor
      buffer = mu
    []
    [Mbar]
      type = ReciprocalLaplacianFactor
      buffer = Mbar
      factor = 0.2
    []
    [Mkappabarbar]
      type = ReciprocalLaplacianSquareFactor
      buffer = Mkappabarbar
      factor = '${fparse 0.2 * 1e-4}'
    []
    [dc_dt_bar_IC]
      type = ConstantReciprocalTensor
      buffer = dc_dt_bar
    []
  []
  [Solve]
    [cahn_hilliard]
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
      [dc_dt_bar]
        type = ParsedCompute
        buffer = dc_dt_bar
        expression = 'Mbar*mubar - Mkappabarbar*cbar'
        inputs = 'Mbar mubar Mkappabarbar cbar'
      []
      [cbar]
        type = ForwardFFT
        buffer = cbar
        input = c
      []
    []
  []
[]

[TensorSolver]
  type = ForwardEulerSolver
  root_compute = cahn_hilliard
  buffer = c
  reciprocal_buffer = cbar
  time_derivative_reciprocal = dc_dt_bar
  substeps = 50
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 100
  dt = 1e-1
[]
