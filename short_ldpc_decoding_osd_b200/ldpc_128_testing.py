"""Drop-in for the NMS test driver ``LDPC_128/Ldpc_128_testing/ldpc_128_testing.py`` as a function.

``main(argv, data_dir, ...)`` follows the reference's module-level script (``:20-156``): for every SNR point read
``test-nonzero<snr>dB-Awgn.tfrecord`` in batches of ``unit_batch_size`` (``:110-111``), decode every batch with
``Decoding_model`` (``:117-131``), stop after ``decoding_threshold`` frame errors (``:36,130``), append the
``FER %.4f, BER %.4f,UFER %.6f`` line to ``./log/FER-<type>-<it>th.txt`` (``:137-140``) and write the 13-row
trajectories of the detected failures to ``ldpc-nonzero-retest.tfrecord`` (``:142-150``).  When a test file is
missing it is generated with the Philox frame generator (the reference expects Testing_data_gen_128 to have run).
"""
from __future__ import annotations

import os

import numpy as np

from . import data_generating as Data_gen
from . import fill_matrix_info as Fill_matrix
from . import globalmap as GL
from . import ms_test as Decoder_module
from . import read_TFdata as Reading

DEFAULT_ARGV = "python 2.0 3.0 6 1000 100 12 CCSDS_ldpc_n128_k64.alist NMS-1".split()


def main(argv=None, data_root="..", frames_if_missing=100000, check_weight=None, log_dir="./log/"):
    argv = list(argv or DEFAULT_ARGV)
    GL.set_map("ALL_ZEROS_CODEWORD_TESTING", False)
    GL.set_map("snr_lo", float(argv[1]))
    GL.set_map("snr_hi", float(argv[2]))
    GL.set_map("snr_num", int(argv[3]))
    GL.set_map("unit_batch_size", int(argv[4]))
    GL.set_map("num_batch_test", int(argv[5]))
    GL.set_map("num_iterations", int(argv[6]))
    GL.set_map("H_filename", argv[7])
    GL.set_map("selected_decoder_type", argv[8])
    GL.set_map("decoding_threshold", 40000)
    GL.set_map("Rayleigh_fading", False)
    decoder_type, n_iteration = argv[8], int(argv[6])
    snr_lo, snr_hi = round(float(argv[1]), 2), round(float(argv[2]), 2)
    path = argv[7] if os.path.exists(argv[7]) else Fill_matrix.CCSDS_ALIST
    code = Fill_matrix.Code(path)
    GL.set_map("code_parameters", code)
    n_dims = code.check_matrix_column
    test_Model = Decoder_module.Decoding_model()
    if check_weight is not None:  # the trained raw weight from the reference's checkpoint (ldpc_128_testing.py:57-68)
        test_Model.layer.shared_check_weight[:] = check_weight
    unit_batch_size = int(argv[4])
    SNRs = np.linspace(snr_lo, snr_hi, int(argv[3]))
    data_dir = os.path.join(data_root, "Testing_data_gen_" + str(n_dims), "data", "snr" + str(snr_lo) + "-" + str(snr_hi) + "dB") + "/"
    os.makedirs(log_dir, exist_ok=True)
    log_filename = log_dir + "FER-" + decoder_type + "-" + str(n_iteration) + "th" + ".txt"
    FER_list = []
    for si, SNR in enumerate(SNRs):
        snr = round(SNR, 2)
        output_dir = data_dir + str(decoder_type) + "/" + str(n_iteration) + "th/" + str(snr) + "dB/"
        os.makedirs(output_dir, exist_ok=True)
        iput_file = data_dir + "test-nonzero" + str(snr) + "dB-Awgn.tfrecord"
        if not os.path.exists(iput_file):
            feats, labels = Data_gen.testing_data_generating(code, SNR, frames_if_missing, seed=si)
            Reading.make_tfrecord((feats, labels), iput_file)
        GL.set_map("noise_standard_variance", np.sqrt(1.0 / (2 * (float(code.k) / float(n_dims)) * 10 ** (SNR / 10))))
        Test_total_fer = Test_total_ber = 0.0
        counter = undetected_sum = 0
        buffer_inputs, buffer_labels = [], []
        for inputs in Reading.data_handler(n_dims, iput_file, unit_batch_size).as_numpy_iterator():
            fer, ber, undetected_count, buffer = test_Model(inputs[0], inputs[1])
            buffer_inputs.append(buffer[0])
            buffer_labels.append(buffer[1])
            Test_total_fer += fer
            Test_total_ber += ber
            undetected_sum += undetected_count
            counter += 1
            if counter % 100 == 0:
                print("%.4f codewords tested, FER:%.4f" % (counter * unit_batch_size, Test_total_fer / counter))
            if Test_total_fer > GL.get_map("decoding_threshold") / unit_batch_size:
                break
        counter = max(counter, 1)
        f1, b1 = Test_total_fer / counter, Test_total_ber / counter
        ufer = undetected_sum / (counter * unit_batch_size)
        print(counter, " batches tested!")
        print("FER %.4f, BER %.4f,UFER %.4f" % (f1, b1, ufer))
        FER_list.append((snr, round(f1, 5)))
        with open(log_filename, "a+") as f:
            f.write("\nFor %.1fdB summary:\n" % round(SNR, 2))
            f.write("FER %.4f, BER %.4f,UFER %.6f" % (f1, b1, ufer) + "\n")
        updated_buffer = test_Model.postprocess_failure_cases((buffer_inputs, buffer_labels))
        Decoder_module.save_decoded_data(updated_buffer, output_dir + "ldpc-nonzero-retest.tfrecord", SNR, log_filename, n_iteration + 1)
    print(f"FER_list:{FER_list}")
    with open(log_filename, "a+") as f:
        f.write(f"FER_list:{FER_list}")
    return FER_list


if __name__ == "__main__":
    main()
