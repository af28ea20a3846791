# Histogram (20 bins on [0, 1]) of the field 0.1 x^2 + 0.2 y^2 + 0.3 z^2 on a 10^3 unit cube.  Same setup as the
# reference's test/tests/histogram/test.i (gold test_out_hist_0001.csv).
[Domain]
  dim = 3
  nx = 10
  ny = 10
  nz = 10
  mesh_mode = DUMMY
[]

[TensorComputes]
  [Initialize]
    [c]
      type = ParsedCompute
      buffer = c
      expression = '0.1*x^2+0.2*y^2+0.3*z^2'
      extra_symbols = true
    []
  []
[]

[VectorPostprocessors]
  [hist]
    type = TensorHistogram
    buffer = c
    bins = 20
    min = 0
    max = 1
    execute_on = 'TIMESTEP_END'
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
  file_base = histogram_out
  execute_on = 'TIMESTEP_END'
[]
