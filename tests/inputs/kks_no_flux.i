# Two-phase KKS-type model inside a smoothed-boundary mask psi (no-flux walls at x = 30 and x = 70 of
# the sampled coordinate): c evolves by variable-mobility diffusion (ReciprocalMatDiffusion), eta by
# Allen-Cahn (ReciprocalAllenCahn + gradient energy), AdamsBashforthMoulton of order 3 with 1000
# substeps.  Same setup as the reference's test/tests/kks/KKS_no_flux_bc.i (gold KKS_no_flux_bc.h5,
# KKS_no_flux_bc_out.csv).
radius = 30
width = 4.2
kappa_eta = 5
rho_sq = 2
w = 1
M = 5
L = 5
ca = 0.3
cb = 0.7

eta0 = '0.5*(1-tanh(2*(sqrt(x^2+y^2)-${radius})/${width}))'
h = 'eta^3*(6*eta^2-15*eta+10)'
F = '${h}*(${rho_sq}*((c - (1-${h})*(${cb} - ${ca}))-${ca})^2) + (1-${h})*(${rho_sq}*((c + (${h})*(${cb} - ${ca}))-${cb})^2 ) + ${w}*(eta^2)*(1-eta)^2'

[Domain]
  dim = 2
  nx = 20
  ny = 20
  xmin = -50
  xmax = 50
  ymin = -50
  ymax = 50
  mesh_mode = DUMMY
[]

[Functions]
  [mask]
    type = ParsedFunction
    expression = 'if(x<x_min-${width},0,if(x>x_min+${width},1,0.5-0.5*cos(pi*(x-(x_min-${width}))/2/${width}) )) * if(x<x_max-${width},1,if(x>x_max+${width},0,0.5+0.5*cos(pi*(x-(x_max-${width}))/2/${width}) ))'
    symbol_names = 'x_min x_max y_min y_max'
    symbol_values = '30 70 0 100'
  []
[]

[TensorComputes]
  [Initialize]
    [c]
      type = ParsedCompute
      buffer = c
      expression = '0.6 + (${ca}-0.6)*${eta0}'
      extra_symbols = true
    []
    [eta]
      type = ParsedCompute
      buffer = eta
      expression = '${eta0}'
      extra_symbols = true
    []
    [psi]
      type = MooseFunctionTensor
      buffer = psi
      function = mask
    []
    [zero]
      type = ConstantReciprocalTensor
      buffer = zero
    []
    [M]
      type = ConstantTensor
      buffer = M
      real = ${M}
    []
    [L]
      type = ConstantTensor
      buffer = L
      real = ${L}
    []
    [L_kappa]
      type = ConstantTensor
      buffer = L_kappa
      real = ${fparse L*kappa_eta}
    []
  []
  [Solve]
    [cbar]
      type = ForwardFFT
      buffer = cbar
      input = c
    []
    [etabar]
      type = ForwardFFT
      buffer = etabar
      input = eta
    []
    [mu]
      type = ParsedCompute
      buffer = mu
      expression = '${F}'
      inputs = 'c eta'
      derivatives = c
    []
    [div_J]
      type = ReciprocalMatDiffusion
      buffer = div_J
      chemical_potential = mu
      mobility = M
      psi = psi
    []
    [domega_deta]
      type = ParsedCompute
      buffer = domega_deta
      expression = '${F} - mu*c'
      inputs = 'mu c eta'
      derivatives = eta
    []
    [AC_bulk]
      type = ReciprocalAllenCahn
      buffer = AC_bulk
      dF_chem_deta = domega_deta
      L = L
      psi = psi
    []
    [kappa_grad_eta]
      type = ReciprocalMatDiffusion
      buffer = kappa_grad_eta
      chemical_potential = eta
      mobility = L_kappa
      psi = psi
    []
    [AC_bar]
      type = ParsedCompute
      buffer = AC_bar
      expression = 'kappa_grad_eta + AC_bulk'
      inputs = 'AC_bulk kappa_grad_eta'
    []
  []
[]

[TensorSolver]
  type = AdamsBashforthMoulton
  buffer = 'c eta'
  reciprocal_buffer = 'cbar etabar'
  linear_reciprocal = 'zero zero'
  nonlinear_reciprocal = 'div_J AC_bar'
  substeps = 1e3
  predictor_order = 3
[]

[Problem]
  type = TensorProblem
[]

[Postprocessors]
  [total_C]
    type = TensorIntegralPostprocessor
    buffer = c
    execute_on = 'INITIAL TIMESTEP_END'
  []
  [total_eta]
    type = TensorIntegralPostprocessor
    buffer = eta
    execute_on = 'INITIAL TIMESTEP_END'
  []
[]

[Executioner]
  type = Transient
  dt = 0.1
  num_steps = 10
[]

[Outputs]
  csv = true
  file_base = kks_no_flux_out
[]
