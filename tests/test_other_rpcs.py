"""The reference client's other three calls (requests.py:67-110) against the in-process servicer: Classify, Regress and
GetModelStatus.  Host-side protobuf only - these are not on the Predict hot path - so the tests run without a GPU."""
import json

import numpy as np
import pytest
from google.protobuf.json_format import MessageToJson

from fake_server import IdentityServer
from min_tfs_client.requests import TensorServingClient, examples_from_input_dict
from tensorflow_serving.apis import classification_pb2, get_model_status_pb2, regression_pb2


@pytest.fixture()
def served():
    srv = IdentityServer()
    yield srv
    srv.stop()


def test_model_status_request_reference_integration_case(served):
    """reference tests/integration/requests_test.py:39-50, verbatim expectations."""
    client = TensorServingClient(host="127.0.0.1", port=served.port, credentials=None)
    response = client.model_status_request(model_name="default")
    response_dict = json.loads(MessageToJson(response))
    assert "model_version_status" in response_dict
    assert len(response_dict["model_version_status"]) == 1
    assert response_dict["model_version_status"][0] == {"version": "1", "state": "AVAILABLE", "status": {}}
    # what travelled: name only (the reference sets the version only when truthy, requests.py:107-108)
    req = get_model_status_pb2.GetModelStatusRequest.FromString(served.received[-1])
    assert req.model_spec.name == "default" and not req.model_spec.HasField("version")
    assert client.model_status_request("default", model_version=7).model_version_status[0].version == 7
    assert get_model_status_pb2.GetModelStatusRequest.FromString(served.received[-1]).model_spec.version.value == 7


def test_examples_from_input_dict():
    inp = examples_from_input_dict({"x": np.array([[1.0, 2.0], [3.0, 4.5]], dtype=np.float32), "id": np.array([7, 8]),
                                    "tag": np.array(["a", "bé"]), "flag": np.bool_(True)})
    ex = inp.example_list.examples
    assert len(ex) == 2
    assert list(ex[1].features.feature["x"].float_list.value) == [3.0, 4.5]
    assert list(ex[0].features.feature["id"].int64_list.value) == [7] and list(ex[1].features.feature["flag"].int64_list.value) == [1]
    assert list(ex[1].features.feature["tag"].bytes_list.value) == ["bé".encode()]
    assert inp.WhichOneof("kind") == "example_list"
    with pytest.raises(ValueError):
        examples_from_input_dict({"a": np.zeros(2), "b": np.zeros(3)})
    with pytest.raises(ValueError):
        examples_from_input_dict({"a": np.array([1 + 2j])})
    assert len(examples_from_input_dict({}).example_list.examples) == 0


def test_classification_and_regression_requests(served):
    client = TensorServingClient(host="127.0.0.1", port=served.port)
    x = np.array([[0.5, 0.25], [2.0, -1.0], [0.0, 0.0]], dtype=np.float32)
    resp = client.classification_request("m", {"x": x, "id": np.arange(3)}, model_version=3)
    assert isinstance(resp, classification_pb2.ClassificationResponse)
    assert [c.classes[0].score for c in resp.result.classifications] == [0.75, 1.0, 0.0]
    assert resp.result.classifications[0].classes[1].label == "negative" and resp.model_spec.version.value == 3
    sent = classification_pb2.ClassificationRequest.FromString(served.received[-1])
    assert sent.model_spec.name == "m" and len(sent.input.example_list.examples) == 3
    resp = client.regression_request("m", {"x": x})
    assert isinstance(resp, regression_pb2.RegressionResponse)
    assert [r.value for r in resp.result.regressions] == [0.75, 1.0, 0.0]
    assert not regression_pb2.RegressionRequest.FromString(served.received[-1]).model_spec.HasField("version")
