# tests/inputs/swift_hohenberg_secant.i with a [TensorSolver/Predictors] block (LinearTensorPredictor on the solver output)
# Swift-Hohenberg (phase-field crystal) model: a rotated hexagonal grain inside a matrix grain, integrated
# with the implicit SecantSolver and an iteration-count driven time step.  Same setup as the reference's
# test/tests/tensor_compute/rotating_grain_secant.i (gold rotating_grain_secant.h5).
w = 6

[Domain]
  dim = 2
  nx = 40
  ny = 40
  xmax = ${fparse w*pi*2}
  ymax = ${fparse w*pi*2/sin(pi/3)}
  mesh_mode = DUMMY
[]

hex = '-(sin(sin(a)*y/2+cos(a)*x/2)^2 + sin(sin(a+1/3*pi)*y/2+cos(a+1/3*pi)*x/2)^2 + sin(sin(a-1/3*pi)*y/2+cos(a-1/3*pi)*x/2)^2 - 1.5)*0.25'

[Functions]
  [matrix]
    type = ParsedFunction
    expression = 'a := 0; ${hex}'
  []
  [grain]
    type = ParsedFunction
    expression = 'a := 0.95; ${hex}'
  []
  [bicrystal]
    type = ParsedFunction
    expression = 'r := (x-${w}*pi)^2+(y-${w}*pi)^2; if(r<(${w}*2/3*pi)^2, grain, matrix)'
    symbol_names = 'matrix grain'
    symbol_values = 'matrix grain'
  []
[]

[TensorComputes]
  [Initialize]
    [psi]
      type = MooseFunctionTensor
      buffer = psi
      function = bicrystal
    []
    [linear]
      type = SwiftHohenbergLinear
      buffer = linear
      alpha = 1
      r = 0.025
    []
  []
  [Solve]
    [psi3]
      type = ParsedCompute
      buffer = psi3
      expression = '0.20*psi^2-psi^3'
      inputs = psi
    []
    [psibar]
      type = ForwardFFT
      buffer = psibar
      input = psi
    []
    [psi3bar]
      type = ForwardFFT
      buffer = psi3bar
      input = psi3
    []
  []
[]

[TensorSolver]
  type = SecantSolver
  buffer = psi
  reciprocal_buffer = psibar
  linear_reciprocal = linear
  nonlinear_reciprocal = psi3bar
  substeps = 3
  [Predictors]
    [linear]
      type = LinearTensorPredictor
      buffer = psi
    []
  []
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 10
  dtmax = 500
  [TimeStepper]
    type = TensorSolveIterationAdaptiveDT
    dt = 1
    min_iterations = 100
    max_iterations = 400
    growth_factor = 1.4
    cutback_factor = 0.9
  []
[]
