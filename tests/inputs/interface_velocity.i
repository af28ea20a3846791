# A sine profile travelling with velocity 0.2, c = sin(x + 0.2 t), re-evaluated every step by a [Solve]
# compute (no solver); TensorInterfaceVelocityPostprocessor recovers the velocity from (c - c_old)/dt and
# the spectral gradient.  Same setup as the reference's test/tests/postprocessors/interface_velocity.i
# (gold interface_velocity_out.csv).
[Domain]
  dim = 2
  nx = 10
  ny = 2
  xmax = '${fparse pi*4}'
  mesh_mode = DUMMY
[]

[TensorComputes]
  [Solve]
    [c]
      type = ParsedCompute
      buffer = c
      expression = 'sin(x+0.2*t)'
      extra_symbols = true
      expand = REAL
    []
  []
[]

[Postprocessors]
  [v]
    type = TensorInterfaceVelocityPostprocessor
    buffer = c
  []
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 10
  dt = 0.01
[]

[Outputs]
  csv = true
  file_base = interface_velocity_out
[]
