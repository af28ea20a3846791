# Quasistatic elastic response to a concentration field with volumetric eigenstrain: per-wavevector 3x3 solve
# for the displacements (FFTQuasistaticElasticity) and the elastic contribution to the chemical potential
# (FFTElasticChemicalPotential).  The operators and parameters are those of the reference's
# test/tests/tensor_compute/group.i / coupled_pf_mech_secant.i (which ship without gold files); the host
# test compares with the oracle.
[Domain]
  dim = 3
  nx = 16
  ny = 12
  nz = 10
  xmax = ${fparse pi*4}
  ymax = ${fparse pi*4}
  zmax = ${fparse pi*4}
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
  []
  [Solve]
    [cbar]
      type = ForwardFFT
      buffer = cbar
      input = c
    []
    [qsmech]
      type = FFTQuasistaticElasticity
      displacements = 'disp_x disp_y disp_z'
      cbar = cbar
      lambda = 100
      mu = 50
      e0 = 0.02
    []
    [mumechbar]
      type = FFTElasticChemicalPotential
      buffer = mumechbar
      cbar = cbar
      displacements = 'disp_x disp_y disp_z'
      lambda = 100
      mu = 50
      e0 = 0.02
    []
    [mumech]
      type = InverseFFT
      buffer = mumech
      input = mumechbar
    []
  []
[]

[Problem]
  type = TensorProblem
[]

[Executioner]
  type = Transient
  num_steps = 1
  dt = 1
[]
