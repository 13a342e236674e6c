"""The reference's integration test (tests/integration/requests_test.py:17-36), run against an
in-process identity server instead of tensorflow_model_server: TensorServingClient.predict_request
end to end - GPU-encoded request over real gRPC, GPU-decoded response - with the reference's API."""
import numpy as np
import pytest
from numpy.testing import assert_array_almost_equal, assert_array_equal

from fake_server import IdentityServer
from min_tfs_client.requests import LARGE_MESSAGE_CHANNEL_OPTIONS, TensorServingClient
from min_tfs_client.tensors import make_ndarray, tensor_proto_to_ndarray
from oracle import ref_port

pytestmark = pytest.mark.gpu


@pytest.fixture()
def served():
    srv = IdentityServer()
    yield srv
    srv.stop()


def test_predict_request_reference_integration_case(served):
    client = TensorServingClient(host="127.0.0.1", port=served.port, credentials=None)
    inputs = {"string_input": np.array(["hello world"]), "float_input": np.array([0.1], dtype=np.float32),
              "int_input": np.array([2], dtype=np.int64)}
    response = client.predict_request(model_name="default", model_version=1, input_dict=inputs)
    assert_array_almost_equal(tensor_proto_to_ndarray(response.outputs["float_output"]), np.array([0.1], dtype=np.float32))
    assert_array_equal(tensor_proto_to_ndarray(response.outputs["int_output"]), np.array([2], dtype=np.int64))
    assert_array_equal(tensor_proto_to_ndarray(response.outputs["string_output"]), np.array(["hello world"]))
    # what travelled is byte for byte what the reference would have sent (deterministic map order)
    assert served.received[-1] == ref_port.encode_predict_request("default", 1, list(inputs.items()), deterministic=True)
    # the lazy view also behaves like the PredictResponse the reference returns
    assert response.model_spec.name == "default" and response.model_spec.version.value == 1
    assert set(response.outputs) == {"float_output", "int_output", "string_output"}
    assert response.outputs["float_output"].dtype == 1 and list(response.outputs["int_output"].int64_val) == [2]


def test_predict_request_large_tensor_and_one_shot_decode(served):
    """A response above grpc's default 4 MiB receive limit needs the channel option the reference never sets
    (SURVEY 8f-3): stay below it here (1 MiB), and decode all outputs in one parse + one unpack."""
    client = TensorServingClient(host="127.0.0.1", port=served.port)
    x = np.random.default_rng(0).standard_normal((512, 512), dtype=np.float32)
    ids = np.arange(-500, 500, dtype=np.int32)
    response = client.predict_request("m", {"x_input": x, "ids_input": ids}, timeout=30)
    outs = response.to_ndarrays(strict=True)
    assert outs["x_output"].tobytes() == x.tobytes() and outs["x_output"].shape == (512, 512)
    assert_array_equal(outs["ids_output"], ids)
    assert_array_equal(make_ndarray(response.outputs["x_output"]), x)
    assert not response.model_spec.HasField("version") or response.model_spec.version.value == 1


def test_c2_sized_round_trip_needs_the_channel_option(served):
    """BASELINE configs[1] over real gRPC: a fp32[1024,1024] response is a few bytes over grpc's default 4 MiB receive
    limit, which the reference never lifts (requests.py:27-30) - same RESOURCE_EXHAUSTED here by default, and the
    round trip, bit-exact, with LARGE_MESSAGE_CHANNEL_OPTIONS."""
    import grpc

    x = np.random.default_rng(1).standard_normal((1024, 1024), dtype=np.float32)
    with pytest.raises(grpc.RpcError) as err:
        TensorServingClient(host="127.0.0.1", port=served.port).predict_request("default", {"x_input": x}, model_version=1, timeout=30)
    assert err.value.code() == grpc.StatusCode.RESOURCE_EXHAUSTED
    client = TensorServingClient(host="127.0.0.1", port=served.port, channel_options=LARGE_MESSAGE_CHANNEL_OPTIONS)
    response = client.predict_request("default", {"x_input": x}, model_version=1, timeout=30)
    assert len(served.received[-1]) == 4194351 + len("_input")          # KAT-3 with the longer key
    assert served.received[-1] == ref_port.encode_predict_request("default", 1, [("x_input", x)], deterministic=True)
    out = tensor_proto_to_ndarray(response.outputs["x_output"])
    assert out.dtype == np.float32 and out.shape == (1024, 1024) and out.tobytes() == x.tobytes()
