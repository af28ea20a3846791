# Postprocessors on a linear ramp c = -x + y + 0.3 over a 40x40 grid of extent 2 x 3: extreme values,
# average, integral, the integral read off the k = 0 mode of the transformed buffer, and the execution
# count of the solver's root compute group.  Same setup as the reference's
# test/tests/postprocessors/postprocessors.i (gold average / integral / extreme_value /
# reciprocal_integral / count .csv).
[Domain]
  dim = 2
  nx = 40
  ny = 40
  xmax = 2
  ymax = 3
  mesh_mode = DUMMY
[]

[TensorComputes]
  [Initialize]
    [c]
      type = ParsedCompute
      buffer = c
      expression = '-x+y+0.3'
      extra_symbols = true
    []
    [c_bar]
      type = ForwardFFT
      buffer = c_bar
      input = c
    []
    [u]
      type = ConstantTensor
      buffer = u
      real = 0
    []
  []
  [Solve]
    [root]
      [test]
        type = ForwardFFT
        buffer = u_bar
        input = u
      []
    []
  []
[]

[TensorSolver]
  type = ForwardEulerSolver
  buffer = u
  reciprocal_buffer = u_bar
  time_derivative_reciprocal = c_bar
  substeps = 10
[]

[Postprocessors]
  [min_c]
    type = TensorExtremeValuePostprocessor
    buffer = c
    value_type = MIN
    execute_on = 'INITIAL TIMESTEP_END'
  []
  [max_c]
    type = TensorExtremeValuePostprocessor
    buffer = c
    value_type = MAX
    execute_on = 'INITIAL TIMESTEP_END'
  []
  [avg_c]
    type = TensorAveragePostprocessor
    buffer = c
    execute_on = 'INITIAL TIMESTEP_END'
  []
  [int_c]
    type = TensorIntegralPostprocessor
    buffer = c
    execute_on = 'INITIAL TIMESTEP_END'
  []
  [int_c_bar]
    type = ReciprocalIntegral
    buffer = c_bar
    execute_on = 'INITIAL TIMESTEP_END'
  []
  [count]
    type = ComputeGroupExecutionCount
  []
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 2
[]

[Outputs]
  csv = true
  file_base = pp_basic
[]
