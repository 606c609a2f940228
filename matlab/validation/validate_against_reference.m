function report = validate_against_reference(ref_task5_dir, gpu_matlab_dir)
%VALIDATE_AGAINST_REFERENCE  Run the UNTOUCHED reference functions and the MEX-backed drop-ins side by side on the same
%   inputs and report the differences (SURVEY 8f rank 4).  To be executed by someone with a MATLAB or Octave host and
%   an sm_100 GPU -- the authoring container has neither interpreter, so this script has not been executed there.
%
%     addpath('<repo>/matlab/validation');
%     report = validate_against_reference('<reference>/Task 5', '<repo>/matlab');
%
%   Prerequisites: `make -C <repo>/ofdm-course_b200` and the gateway built with `mex` / `mkoctfile --mex`
%   (INTEGRATION.md), both on the path via gpu_matlab_dir.  Integer outputs must match exactly; complex outputs are
%   compared in relative L2 norm against the FP32 tolerances of DESIGN.md section 2 (call ofdm_mex('precision','f64')
%   first for the FP64 comparison mode).
    install_shims();
    names = {'Scrambler', 'DeScrambler', 'constellation_func', 'mapping', 'demapping', 'OFDM_map_carriers', 'OFDM_modulator', ...
             'OFDM_demodulator', 'get_payload', 'add_STO', 'add_CFO', 'get_MP_channel_resp', 'AutoCorrFunction', 'remove_IFO', ...
             'LS_CE', 'MMSE_CE', 'interpolate', 'equalize_signal', 'OMP_estimate', 'MP_estimate', 'BER_func', 'MER_func', ...
             'estimate_channel', 'calculatePAPR', 'calculate_window_PAPR', 'calculateCCDF'};
    R = handles_from(ref_task5_dir, names);
    G = handles_from(gpu_matlab_dir, names);
    rng(1);
    report = struct('name', {}, 'err', {}, 'tol', {}, 'ok', {});
    function add(name, err, tol)
        report(end + 1) = struct('name', name, 'err', err, 'tol', tol, 'ok', err <= tol); %#ok<AGROW>
        fprintf('%-28s err %-12.3g tol %-9.3g %s\n', name, err, tol, ternary(err <= tol, 'ok', 'MISMATCH'));
    end
    rel = @(a, b) norm(a(:) - b(:)) / max(norm(b(:)), realmin);

    % ---- Task-5 part-2 shape, comb 4
    Nfft = 4096; Nc = 1024; Tg = Nfft / 8; S = 14; comb = 4;
    pilots = 1:comb:Nc; data = setdiff(1:Nc, pilots);
    [dict, bps] = R.constellation_func("16QAM");
    [dict_g, bps_g] = G.constellation_func("16QAM");
    add('constellation_func', max(abs(dict(:) - dict_g(:))) + abs(bps - bps_g), 1e-15);
    amp = 2 * max(abs(dict));
    pv = zeros(1, numel(pilots)); pv(1:2:end) = amp; pv(2:2:end) = amp * exp(1i * pi); pv = repmat(pv', 1, S);
    bits = double(rand(1, S * numel(data) * bps) > 0.5);
    reg = [1 0 0 1 0 1 0 1 0 0 0 0 0 0 0];
    frame = numel(bits) / 2;

    [s_r, r_r] = R.Scrambler(reg, bits(1:frame)); [s_g, r_g] = G.Scrambler(reg, bits(1:frame));
    add('Scrambler', sum(s_r ~= s_g) + sum(r_r ~= r_g), 0);
    d_r = R.DeScrambler(reg, s_r); d_g = G.DeScrambler(reg, s_r);
    add('DeScrambler', sum(d_r ~= d_g) + sum(d_r ~= bits(1:frame)), 0);
    sc = [R.Scrambler(reg, bits(1:frame)), R.Scrambler(reg, bits(frame + 1:end))];
    [iq_r, pad_r] = R.mapping(sc.', "16QAM"); [iq_g, pad_g] = G.mapping(sc.', "16QAM");
    add('mapping', rel(iq_g, iq_r) + abs(pad_r - pad_g), 1e-7);
    grid_r = R.OFDM_map_carriers(iq_r, S, Nfft, data, pilots, pv); grid_g = G.OFDM_map_carriers(iq_r, S, Nfft, data, pilots, pv);
    add('OFDM_map_carriers', rel(grid_g, grid_r), 1e-7);
    tx_r = R.OFDM_modulator(grid_r, Tg); tx_g = G.OFDM_modulator(grid_r, Tg);
    add('OFDM_modulator', rel(tx_g, tx_r), 2e-6);
    add('calculatePAPR', abs(G.calculatePAPR(tx_r(:)) - R.calculatePAPR(tx_r(:))), 1e-4);
    w_r = R.calculate_window_PAPR(tx_r(1:3 * Nfft).', Nfft); w_g = G.calculate_window_PAPR(tx_r(1:3 * Nfft).', Nfft);
    add('calculate_window_PAPR', max(abs(w_r(:) - w_g(:))), 1e-3);
    if exist('ecdf', 'file')
        [x_r, c_r] = R.calculateCCDF(round(w_r, 2)); [x_g, c_g] = G.calculateCCDF(round(w_r, 2));
        add('calculateCCDF', double(numel(x_r) ~= numel(x_g)) + max(abs(x_r(:) - x_g(1:numel(x_r)))) + max(abs(c_r(:) - c_g(1:numel(c_r)))), 1e-5);
    end

    % ---- channel: the reference's order, noise first (Main_model_Task_5.m:108,123-127); the normals are drawn here
    %      in the order Noise.m draws them and handed to the GPU side as the optional third argument
    x = tx_r(:);
    n1 = normrnd(0, 1, size(x)); n2 = normrnd(0, 1, size(x));
    P = mean(abs(x) .^ 2) / 10 ^ (20 / 10);
    rx_ref_noise = x + sqrt(P / 2) * n1 + 1i * sqrt(P / 2) * n2;
    rx_g = ofdm_mex('Noise', 20, x, [n1, n2]);
    add('Noise (imported normals)', rel(rx_g, rx_ref_noise), 2e-6);
    taps = [0 1; 4 .8; 10 .6; 15 .4; 21 .2; 25 .1];
    [h_r, H_r] = R.get_MP_channel_resp(taps, Nfft); [h_g, H_g] = G.get_MP_channel_resp(taps, Nfft);
    add('get_MP_channel_resp', rel(h_g, h_r) + rel(H_g, H_r), 1e-6);
    y = conv(rx_ref_noise, h_r.', 'full'); y = y(1:numel(x));
    add('add_STO', rel(G.add_STO(y, 37), R.add_STO(y, 37)) + rel(G.add_STO(y, -37), R.add_STO(y, -37)), 0);
    add('add_CFO', rel(G.add_CFO(y, 7.24, Nfft), R.add_CFO(y, 7.24, Nfft)), 2e-6);

    % ---- RX: demodulator, estimators, equaliser, demapper
    Y_r = R.OFDM_demodulator(reshape(y, Nfft + Tg, S), Tg); Y_g = G.OFDM_demodulator(reshape(y, Nfft + Tg, S), Tg);
    add('OFDM_demodulator', rel(Y_g, Y_r), 2e-5);
    H_ls_r = R.LS_CE(Y_r, pv, pilots, Nc); H_ls_g = G.LS_CE(Y_r, pv, pilots, Nc);
    add('LS_CE', rel(H_ls_g, H_ls_r), 2e-5);
    Hp = Y_r(pilots, 1) ./ pv(:, 1);
    add('interpolate (spline)', rel(G.interpolate(Hp.', pilots, Nc, 'spline'), R.interpolate(Hp.', pilots, Nc, 'spline')), 2e-5);
    add('interpolate (linear)', rel(G.interpolate(Hp.', pilots, Nc, 'linear'), R.interpolate(Hp.', pilots, Nc, 'linear')), 2e-5);
    h_true = [h_r, zeros(1, Nc - numel(h_r))];
    add('MMSE_CE', rel(G.MMSE_CE(Y_r, pv, pilots, Nfft, Nc, h_true, 20), R.MMSE_CE(Y_r, pv, pilots, Nfft, Nc, h_true, 20)), 5e-5);
    F = dftmtx(Nfft); F = F(:, 1:ceil(Nfft / comb)); A = F(pilots, :);
    [Ho_r, ho_r, ix_r] = R.OMP_estimate(Hp, A, Nfft, 6, 20); [Ho_g, ho_g, ix_g] = G.OMP_estimate(Hp, A, Nfft, 6, 20);
    add('OMP_estimate (indices)', double(~isequal(ix_r(:), ix_g(:))), 0);
    add('OMP_estimate (H, h)', rel(Ho_g, Ho_r) + rel(ho_g, ho_r), 2e-4);
    [Hm_r, hm_r] = R.MP_estimate(Hp, A, Nfft, 6); [Hm_g, hm_g] = G.MP_estimate(Hp, A, Nfft, 6);
    add('MP_estimate', rel(Hm_g, Hm_r) + rel(hm_g, hm_r), 2e-4);
    eq_r = R.equalize_signal(Y_r, H_ls_r, Nc); eq_g = G.equalize_signal(Y_r, H_ls_r, Nc);
    add('equalize_signal', rel(eq_g, eq_r), 2e-6);
    p_r = R.get_payload(eq_r, data); p_g = G.get_payload(eq_r, data);
    add('get_payload', rel(p_g, p_r), 0);
    b_r = R.demapping(pad_r, p_r(:).', "16QAM"); b_g = G.demapping(pad_r, p_r(:).', "16QAM");
    near = sum(min(abs(abs(real(p_r(:))) - [0, 2 / sqrt(10)]), [], 2) < 1e-5 | min(abs(abs(imag(p_r(:))) - [0, 2 / sqrt(10)]), [], 2) < 1e-5);
    add('demapping (bit mismatches)', sum(b_r ~= b_g), 4 * near);
    add('BER_func', abs(G.BER_func(sc, b_r) - R.BER_func(sc, b_r)), 0);
    add('MER_func', abs(G.MER_func(p_r(:).', "16QAM") - R.MER_func(p_r(:).', "16QAM")), 1e-3);

    % ---- Task-4 shape: synchronisation functions
    N4 = 1024; T4 = 128; S4 = 50; Nc4 = 400;
    pil4 = [1:6:398, 400]; dat4 = setdiff(1:Nc4, pil4);
    a4 = 4 / 3 * max(abs(dict)); pv4 = zeros(1, numel(pil4)); pv4(1:2:end) = a4; pv4(2:2:end) = a4 * exp(1i * pi); pv4 = repmat(pv4', 1, S4);
    b4 = double(rand(S4 * numel(dat4) * bps, 1) > 0.5);
    g4 = R.OFDM_map_carriers(R.mapping(b4, "16QAM"), S4, N4, dat4, pil4, pv4);
    t4 = R.OFDM_modulator(g4, T4); t4 = t4(:);
    r4 = R.add_CFO(R.add_STO(t4, 611), 3.3, N4);
    [ac_r, tg_r, fo_r] = R.AutoCorrFunction(r4, T4, N4); [ac_g, tg_g, fo_g] = G.AutoCorrFunction(r4, T4, N4);
    ok = isfinite(ac_r(:));
    add('AutoCorrFunction (TgPosition)', abs(tg_r - tg_g), 0);
    add('AutoCorrFunction (rho, FreqOffset)', rel(ac_g(ok), ac_r(ok)) + abs(fo_r - fo_g), 2e-4);
    y3 = R.add_CFO(R.add_STO(R.add_STO(r4, tg_r), -(N4 + T4)), -fo_r, N4);
    [f_r, i_r] = R.remove_IFO(y3, N4); [f_g, i_g] = G.remove_IFO(y3, N4);
    add('remove_IFO', abs(i_r - i_g) + rel(f_g, f_r), 2e-5);
    Y4 = R.OFDM_demodulator(reshape(f_r, N4 + T4, S4), T4);
    [He_r, Hp_r] = R.estimate_channel(Y4, 1:N4, pil4, pv4); [He_g, Hp_g] = G.estimate_channel(Y4, 1:N4, pil4, pv4);
    add('estimate_channel', rel(He_g(1:Nc4), He_r(1:Nc4)) + rel(Hp_g, Hp_r), 1e-4);
    if exist(fullfile(ref_task5_dir, '..', 'Task 4', 'fine_sync.m'), 'file')
        R4 = handles_from(fullfile(ref_task5_dir, '..', 'Task 4'), {'fine_sync'});
        G4 = handles_from(gpu_matlab_dir, {'fine_sync'});
        add('fine_sync', rel(G4.fine_sync(Y4, pil4, pv4, 1, 1), R4.fine_sync(Y4, pil4, pv4, 1, 1)), 2e-4);
    end
    fprintf('%d of %d checks ok\n', sum([report.ok]), numel(report));
end

function v = ternary(c, a, b)
    if c, v = a; else, v = b; end
end
