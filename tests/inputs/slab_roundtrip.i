# 2-D round trip through the slab-decomposed transforms on three ranks (restates the check of the reference's
# test/tests/tensor_compute/parallel_roundtrip.i with the transforms in [Initialize])
[Domain]
  device_names = "cuda cuda cuda"
  device_weights = "1 1 1"
  parallel_mode = FFT_SLAB
  dim = 2
  nx = 128
  ny = 128
  xmax = ${fparse pi*4}
  ymax = ${fparse pi*4}
[]

[TensorBuffers]
  [eta_gold]
  []
  [eta]
  []
  [eta_bar]
  []
  [eta_roundtrip]
  []
  [diff]
  []
[]

[TensorComputes]
  [Initialize]
    [eta_gold]
      type = ParsedCompute
      buffer = eta_gold
      expression = 'sin(x)+sin(y)+cos(2*x)*sin(3*y)'
      extra_symbols = true
    []
    [eta]
      type = ParsedCompute
      buffer = eta
      expression = eta_gold
      inputs = eta_gold
    []
    [eta_bar]
      type = ForwardFFT
      buffer = eta_bar
      input = eta
    []
    [eta_roundtrip]
      type = InverseFFT
      buffer = eta_roundtrip
      input = eta_bar
    []
  []
  [Postprocess]
    [diff]
      type = ParsedCompute
      buffer = diff
      expression = 'abs(eta - eta_roundtrip) + abs(eta - eta_gold)'
      inputs = 'eta eta_roundtrip eta_gold'
    []
  []
[]

[Postprocessors]
  [max_error]
    type = TensorExtremeValuePostprocessor
    buffer = diff
    value_type = MAX
  []
  [l2_error]
    type = TensorIntegralPostprocessor
    buffer = diff
  []
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 1
[]

[Outputs]
  csv = true
  execute_on = 'INITIAL TIMESTEP_END'
[]
