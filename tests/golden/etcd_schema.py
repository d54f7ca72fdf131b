"""The etcd v3.5.2 message schemas the reference's etcd-compatible server answers with (go.etcd.io/etcd/api/v3 v3.5.2,
go.mod:28 of the reference; the module is not vendored).  Restated from the published kv.proto / rpc.proto field
numbers and built with the protobuf runtime, so the serialiser that produces the golden bytes is the official one."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _field(m, name, number, ftype, label=_F.LABEL_OPTIONAL, type_name=None):
    f = m.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name


def build():
    fdp = descriptor_pb2.FileDescriptorProto(name="etcd_restated.proto", package="etcdr", syntax="proto3")
    kv = fdp.message_type.add(name="KeyValue")  # mvccpb.KeyValue
    _field(kv, "key", 1, _F.TYPE_BYTES)
    _field(kv, "create_revision", 2, _F.TYPE_INT64)
    _field(kv, "mod_revision", 3, _F.TYPE_INT64)
    _field(kv, "version", 4, _F.TYPE_INT64)
    _field(kv, "value", 5, _F.TYPE_BYTES)
    _field(kv, "lease", 6, _F.TYPE_INT64)
    ev = fdp.message_type.add(name="Event")  # mvccpb.Event (type: PUT=0 / DELETE=1)
    _field(ev, "type", 1, _F.TYPE_INT32)
    _field(ev, "kv", 2, _F.TYPE_MESSAGE, type_name=".etcdr.KeyValue")
    _field(ev, "prev_kv", 3, _F.TYPE_MESSAGE, type_name=".etcdr.KeyValue")
    hd = fdp.message_type.add(name="ResponseHeader")  # etcdserverpb.ResponseHeader
    _field(hd, "cluster_id", 1, _F.TYPE_UINT64)
    _field(hd, "member_id", 2, _F.TYPE_UINT64)
    _field(hd, "revision", 3, _F.TYPE_INT64)
    _field(hd, "raft_term", 4, _F.TYPE_UINT64)
    rr = fdp.message_type.add(name="RangeResponse")  # etcdserverpb.RangeResponse
    _field(rr, "header", 1, _F.TYPE_MESSAGE, type_name=".etcdr.ResponseHeader")
    _field(rr, "kvs", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".etcdr.KeyValue")
    _field(rr, "more", 3, _F.TYPE_BOOL)
    _field(rr, "count", 4, _F.TYPE_INT64)
    wr = fdp.message_type.add(name="WatchResponse")  # etcdserverpb.WatchResponse
    _field(wr, "header", 1, _F.TYPE_MESSAGE, type_name=".etcdr.ResponseHeader")
    _field(wr, "watch_id", 2, _F.TYPE_INT64)
    _field(wr, "created", 3, _F.TYPE_BOOL)
    _field(wr, "canceled", 4, _F.TYPE_BOOL)
    _field(wr, "compact_revision", 5, _F.TYPE_INT64)
    _field(wr, "cancel_reason", 6, _F.TYPE_STRING)
    _field(wr, "fragment", 7, _F.TYPE_BOOL)
    _field(wr, "events", 11, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".etcdr.Event")
    pool = descriptor_pool.DescriptorPool()
    fd = pool.Add(fdp) if hasattr(pool, "Add") and not hasattr(pool, "AddSerializedFile") else pool.AddSerializedFile(
        fdp.SerializeToString())
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("etcdr." + n))
    return {n: get(n) for n in ("KeyValue", "Event", "ResponseHeader", "RangeResponse", "WatchResponse")}


def s64(u: int) -> int:
    """uint64 -> the int64 Go's int64(kv.Revision) yields (backendshim.go:434)"""
    return u - (1 << 64) if u >= (1 << 63) else u


def range_response(M, header_rev: int, kvs, more: bool, count: int) -> bytes:
    """backendshim.go:269-282: kvs = [(key, value, revision)]"""
    r = M["RangeResponse"]()
    r.header.SetInParent()
    r.header.revision = s64(header_rev)
    for k, v, rev in kvs:
        r.kvs.add(key=k, value=v, mod_revision=s64(rev))
    r.more, r.count = more, count
    return r.SerializeToString(deterministic=True)


def watch_batch(M, header_rev: int, kvs) -> bytes:
    """backendshim.go:349-363 (More == true): one WatchResponse per StreamRangeResponse batch"""
    r = M["WatchResponse"]()
    r.header.SetInParent()
    r.header.revision = s64(header_rev)
    for k, v, rev in kvs:
        e = r.events.add()
        e.kv.SetInParent()
        e.kv.key, e.kv.value, e.kv.mod_revision = k, v, s64(rev)
    return r.SerializeToString(deterministic=True)


def watch_cancel(M, header_rev: int, reason: str) -> bytes:
    """backendshim.go:353-355 (More == false): the end-of-stream / error message"""
    r = M["WatchResponse"]()
    r.header.SetInParent()
    r.header.revision = s64(header_rev)
    r.canceled, r.cancel_reason = True, reason
    return r.SerializeToString(deterministic=True)
