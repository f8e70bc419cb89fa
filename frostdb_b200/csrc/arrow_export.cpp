#include "arrow_export.h"

#include <cstdlib>
#include <cstring>

namespace fgpu {
namespace {

struct SchemaPriv {
  std::string format, name;
  std::vector<ArrowSchema*> children;
  ArrowSchema* dictionary = nullptr;
};

void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  auto* p = static_cast<SchemaPriv*>(s->private_data);
  for (ArrowSchema* c : p->children) {
    if (c->release) c->release(c);
    delete c;
  }
  if (p->dictionary) {
    if (p->dictionary->release) p->dictionary->release(p->dictionary);
    delete p->dictionary;
  }
  delete p;
  s->release = nullptr;
}

struct ArrayPriv {
  OwnedColumn col;  // owns the buffers
  std::vector<const void*> buffers;
  std::vector<ArrowArray*> children;
  ArrowArray* dictionary = nullptr;
};

void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  auto* p = static_cast<ArrayPriv*>(a->private_data);
  for (ArrowArray* c : p->children) {
    if (c->release) c->release(c);
    delete c;
  }
  if (p->dictionary) {
    if (p->dictionary->release) p->dictionary->release(p->dictionary);
    delete p->dictionary;
  }
  delete p;
  a->release = nullptr;
}

void fill_schema(const OwnedColumn& col, ArrowSchema* s) {
  auto* p = new SchemaPriv;
  p->format = col.format;
  p->name = col.name;
  std::memset(s, 0, sizeof(*s));
  s->flags = ARROW_FLAG_NULLABLE;
  if (col.dictionary) {
    p->dictionary = new ArrowSchema;
    fill_schema(*col.dictionary, p->dictionary);
  }
  s->format = p->format.c_str();
  s->name = p->name.c_str();
  s->dictionary = p->dictionary;
  s->release = release_schema;
  s->private_data = p;
}

void fill_array(OwnedColumn&& col, ArrowArray* a) {
  auto* p = new ArrayPriv;
  p->col = std::move(col);
  OwnedColumn& c = p->col;
  std::memset(a, 0, sizeof(*a));
  a->length = c.length;
  a->null_count = c.null_count;
  const void* validity = c.validity.empty() ? nullptr : c.validity.data();
  // Zero-length buffers must still be non-null pointers for some consumers.
  static const uint64_t kEmpty[2] = {0, 0};
  if (c.format == "z" || c.format == "u") {
    p->buffers = {validity, c.offsets.empty() ? static_cast<const void*>(kEmpty) : c.offsets.data(),
                  c.data.empty() ? static_cast<const void*>(kEmpty) : c.data.data()};
  } else {
    p->buffers = {validity, c.ext ? static_cast<const void*>(c.ext) : (c.data.empty() ? static_cast<const void*>(kEmpty) : c.data.data())};
  }
  if (c.dictionary) {
    p->dictionary = new ArrowArray;
    fill_array(std::move(*c.dictionary), p->dictionary);
    c.dictionary.reset();
  }
  a->n_buffers = int64_t(p->buffers.size());
  a->buffers = p->buffers.data();
  a->dictionary = p->dictionary;
  a->release = release_array;
  a->private_data = p;
}

}  // namespace

void export_column(OwnedColumn&& col, ArrowSchema* out_schema, ArrowArray* out_array) {
  fill_schema(col, out_schema);
  fill_array(std::move(col), out_array);
}

void export_record(std::vector<OwnedColumn>&& cols, int64_t length, ArrowSchema* out_schema, ArrowArray* out_array) {
  auto* sp = new SchemaPriv;
  sp->format = "+s";
  sp->name = "";
  auto* ap = new ArrayPriv;
  for (auto& c : cols) {
    auto* cs = new ArrowSchema;
    fill_schema(c, cs);
    sp->children.push_back(cs);
    auto* ca = new ArrowArray;
    fill_array(std::move(c), ca);
    ap->children.push_back(ca);
  }
  std::memset(out_schema, 0, sizeof(*out_schema));
  out_schema->format = sp->format.c_str();
  out_schema->name = sp->name.c_str();
  out_schema->flags = 0;
  out_schema->n_children = int64_t(sp->children.size());
  out_schema->children = sp->children.data();
  out_schema->release = release_schema;
  out_schema->private_data = sp;

  std::memset(out_array, 0, sizeof(*out_array));
  ap->buffers = {nullptr};
  out_array->length = length;
  out_array->null_count = 0;
  out_array->n_buffers = 1;
  out_array->buffers = ap->buffers.data();
  out_array->n_children = int64_t(ap->children.size());
  out_array->children = ap->children.data();
  out_array->release = release_array;
  out_array->private_data = ap;
}

}  // namespace fgpu
