function export_golden_realisation(ref_task5_dir, out_file)
%EXPORT_GOLDEN_REALISATION  Produce tests/golden/realisation_small.mat from the UNTOUCHED reference functions: the
%   same small shape, variable names and layout as tests/golden/make_realisation_fixture.py, so that replacing the
%   committed file with this one pins the oracle and the GPU path to MATLAB's own numbers (run
%   `pytest tests/test_realisations.py tests/test_gpu_realisation.py` afterwards).  Not executed in the authoring
%   container (no MATLAB/Octave there).
%
%     export_golden_realisation('<reference>/Task 5', '<repo>/tests/golden/realisation_small.mat')
    install_shims();
    R = handles_from(ref_task5_dir, {'Scrambler', 'DeScrambler', 'constellation_func', 'mapping', 'demapping', 'OFDM_map_carriers', ...
                                     'OFDM_modulator', 'OFDM_demodulator', 'get_payload', 'get_MP_channel_resp', 'LS_CE', 'equalize_signal'});
    Nfft = 512; Nc = 128; Tg = 64; frames = 2; spf = 2; S = frames * spf; comb = 4; SNR_dB = 9;
    pilots = 1:comb:Nc; data = setdiff(1:Nc, pilots);
    [dict, bps] = R.constellation_func("16QAM");
    amp = 2 * max(abs(dict));
    pv = zeros(1, numel(pilots)); pv(1:2:end) = amp * exp(1i * 0); pv(2:2:end) = amp * exp(1i * pi); pv = repmat(pv', 1, S);
    rng(20261018);
    input_bits = double(rand(1, S * numel(data) * bps) > 0.5);
    reg = [1 0 0 1 0 1 0 1 0 0 0 0 0 0 0];
    fb = numel(input_bits) / frames;
    sc = zeros(1, numel(input_bits));
    for f = 1:frames, sc((f - 1) * fb + 1:f * fb) = R.Scrambler(reg, input_bits((f - 1) * fb + 1:f * fb)); end
    IQ = R.mapping(sc.', "16QAM");
    tx = R.OFDM_modulator(R.OFDM_map_carriers(IQ, S, Nfft, data, pilots, pv), Tg);
    ref_tx = tx(:);
    n1 = normrnd(0, 1, size(ref_tx)); n2 = normrnd(0, 1, size(ref_tx));          % the two draws of Noise.m:7-8, in its order
    noise_normals = [n1, n2];
    P = mean(abs(ref_tx) .^ 2) / 10 ^ (SNR_dB / 10);
    noisy = ref_tx + sqrt(P / 2) * n1 + 1i * sqrt(P / 2) * n2;
    channel_taps = [0 1; 4 .8; 10 .6];
    h = R.get_MP_channel_resp(channel_taps, Nfft);
    ref_rx = conv(noisy, h.', 'full'); ref_rx = ref_rx(1:numel(ref_tx));         % Main_model_Task_5.m:126-127
    Y = R.OFDM_demodulator(reshape(ref_rx, Nfft + Tg, S), Tg);
    ref_H_LS = R.LS_CE(Y, pv, pilots, Nc);
    eq = R.equalize_signal(Y, ref_H_LS, Nc);
    p = R.get_payload(eq, data);
    raw = R.demapping(-1, p(:).', "16QAM");
    ref_bits = zeros(1, numel(raw));
    for f = 1:frames, ref_bits((f - 1) * fb + 1:f * fb) = R.DeScrambler(reg, raw((f - 1) * fb + 1:f * fb)); end
    ref_errors = sum(ref_bits ~= input_bits);
    ref_tx = ref_tx.'; ref_rx = ref_rx.'; %#ok<NASGU>
    save(out_file, 'input_bits', 'noise_normals', 'channel_taps', 'SNR_dB', 'ref_tx', 'ref_rx', 'ref_H_LS', 'ref_bits', 'ref_errors', '-v6');
    fprintf('wrote %s (%d bit errors of %d)\n', out_file, ref_errors, numel(input_bits));
end
