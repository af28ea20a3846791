# A box [5,15]^2 in a 100^2 grid on [0,20]^2, value -1 inside and 3 outside, with a sharp, a half-sine and a tanh
# interface.  Same setup as the reference's test/tests/tensor_compute/smooth_rectangle.i (gold smooth_rectangle.h5);
# one (empty) time step instead of none so that the TIMESTEP_END path of the driver is the one used.
[Domain]
  dim = 2
  nx = 100
  ny = 100
  xmax = 20
  ymax = 20
  mesh_mode = DUMMY
[]

[TensorComputes]
  [Initialize]
    [rectangle_sharp]
      type = SmoothRectangleCompute
      buffer = rectangle_sharp
      x1 = 5
      x2 = 15
      y1 = 5
      y2 = 15
      inside = -1
      outside = 3
    []
    [rectangle_cos]
      type = SmoothRectangleCompute
      buffer = rectangle_cos
      x1 = 5
      x2 = 15
      y1 = 5
      y2 = 15
      inside = -1
      outside = 3
      profile = COS
      int_width = 1
    []
    [rectangle_tanh]
      type = SmoothRectangleCompute
      buffer = rectangle_tanh
      x1 = 5
      x2 = 15
      y1 = 5
      y2 = 15
      inside = -1
      outside = 3
      profile = TANH
      int_width = 1
    []
  []
[]

[Problem]
  type = TensorProblem
[]

[TensorOutputs]
  active = ''                      # TensorOutputs/active=xdmf writes the .xmf + raw data files
  [xdmf]
    type = XDMFTensorOutput
    buffer = 'rectangle_sharp rectangle_cos rectangle_tanh'
  []
[]

[Executioner]
  type = Transient
  num_steps = 1
[]
