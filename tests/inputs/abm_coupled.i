# Two diffusing fields coupled through an off-diagonal LINEAR operator (dense 2x2 solve per
# wavevector): AdamsBashforthMoultonCoupled of order ${order}, ${cs} corrector steps, ${ss} substeps
# (command line: ss=10 cs=0 order=2).  Same setup as the reference's test/tests/solvers/coupled.i,
# whose gold CSVs (coupled_<ss>_<cs>_<order>.csv) the host test compares with.
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
  []
[]

[TensorSolver]
  type = AdamsBashforthMoultonCoupled
  buffer = 'u v'
  reciprocal_buffer = 'u_bar v_bar'
  linear_reciprocal = 'D1 D1'
  linear_offdiag_rows = '1 0'
  linear_offdiag_cols = '0 1'
  linear_offdiag = 'D2 D2'
  nonlinear_reciprocal = 'zero zero'
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
  file_base = abm_coupled_${ss}_${cs}_${order}
  csv = true
[]
