# The legacy per-buffer integrator FFTSemiImplicit (src/tensor_timeintegrators/FFTSemiImplicit.C:43-62) used as a
# TensorOperator inside the root compute of a ForwardEulerSolver that integrates nothing and only forwards
# cnew -> c (the pattern of test/tests/mechanics/mech3d.i): the solver supplies the sub step, FFTSemiImplicit
# the semi-implicit update (first order while there is no history - MOOSE step 1, quirk Q1 - then the
# two-level Adams-Bashforth combination).
[Domain]
  dim = 2
  nx = 32
  ny = 24
  xmax = 4
  ymax = 3
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
      factor = -0.001
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
    [cnew]
      type = FFTSemiImplicit
      buffer = cnew
      reciprocal_buffer = cbar
      linear_reciprocal = kappabarbar
      nonlinear_reciprocal = Mbarmubar
    []
  []
[]

[TensorSolver]
  type = ForwardEulerSolver
  substeps = 5
  forward_buffer = c
  forward_buffer_new = cnew
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 3
  dt = 5e-3
[]
