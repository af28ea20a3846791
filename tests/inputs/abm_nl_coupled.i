# Two diffusing fields with the cross coupling written as complex-valued NONLINEAR terms
# (D2*v_bar, D2*u_bar) of the diagonal AdamsBashforthMoulton solver: order ${order}, ${cs} corrector
# steps, ${ss} substeps.  Same setup as the reference's test/tests/solvers/nl_coupled.i
# (gold nl_coupled_<ss>_<cs>_<order>.csv).
[Domain]
  dim = 2
  nx = 150
  ny = 150
  xmax = '${fparse pi*2}'
  ymax = '${fparse pi*2}'
  mesh_mode = DUMMY
[]

[TensorComputes]
  [Initialize]
    [u]
      type = ParsedCompute
      buffer = u
      expression = 'sin(x)*sin(y)'
      extra_symbols = true
      expand = REAL
    []
    [v]
      type = ParsedCompute
      buffer = v
      expression = 'cos(x)*cos(y)'
      extra_symbols = true
      expand = REAL
    []
    [zero]
      type = ConstantReciprocalTensor
      buffer = zero
    []
    [D1]
      type = ReciprocalLaplacianFactor
      buffer = D1
      factor = 1e-2
    []
    [D2]
      type = ReciprocalLaplacianFactor
      buffer = D2
      factor = 1e-3
    []
  []
  [Solve]
    [u_bar]
      type = ForwardFFT
      buffer = u_bar
      input = u
    []
    [v_bar]
      type = ForwardFFT
      buffer = v_bar
      input = v
    []
    [Du]
      type = ParsedCompute
      buffer = Du
      expression = 'D2*v_bar'
      inputs = 'D2 v_bar'
    []
    [Dv]
      type = ParsedCompute
      buffer = Dv
      expression = 'D2*u_bar'
      inputs = 'D2 u_bar'
    []
  []
[]

[TensorSolver]
  type = AdamsBashforthMoulton
  buffer = 'u v'
  reciprocal_buffer = 'u_bar v_bar'
  linear_reciprocal = 'D1 D1'
  nonlinear_reciprocal = 'Du Dv'
  substeps = ${ss}
  predictor_order = ${order}
  corrector_order = ${order}
  corrector_steps = ${cs}
[]

[Problem]
  type = TensorProblem
[]

[Postprocessors]
  [U]
    type = TensorIntegralPostprocessor
    buffer = u
  []
  [V]
    type = TensorIntegralPostprocessor
    buffer = v
  []
  [u_max]
    type = TensorExtremeValuePostprocessor
    buffer = u
    value_type = MAX
  []
  [u_min]
    type = TensorExtremeValuePostprocessor
    buffer = u
    value_type = MIN
  []
  [v_max]
    type = TensorExtremeValuePostprocessor
    buffer = v
    value_type = MAX
  []
  [v_min]
    type = TensorExtremeValuePostprocessor
    buffer = v
    value_type = MIN
  []
[]

[Executioner]
  type = Transient
  num_steps = 25
  dt = 10
[]

[Outputs]
  file_base = abm_nl_coupled_${ss}_${cs}_${order}
  csv = true
[]
