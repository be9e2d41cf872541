/* cr_json.c — recursive-descent JSON reader (see cr_json.h). */
#include "cr_json.h"
#include <ctype.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

struct parser { const char *p; bool err; };

static void skip_ws(struct parser *s) { while (*s->p && (unsigned char)*s->p <= 32) s->p++; }
static struct crj *parse_value(struct parser *s, int depth);

static struct crj *node_new(enum crj_type t) {
	struct crj *n = calloc(1, sizeof *n);
	if (n) n->type = t;
	return n;
}

void crj_free(struct crj *j) {
	while (j) {
		struct crj *next = j->next;
		crj_free(j->child);
		free(j->key); free(j->str); free(j);
		j = next;
	}
}

static void utf8_put(char **out, unsigned cp) {
	char *o = *out;
	if (cp < 0x80) *o++ = (char)cp;
	else if (cp < 0x800) { *o++ = (char)(0xC0 | (cp >> 6)); *o++ = (char)(0x80 | (cp & 0x3F)); }
	else if (cp < 0x10000) { *o++ = (char)(0xE0 | (cp >> 12)); *o++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *o++ = (char)(0x80 | (cp & 0x3F)); }
	else { *o++ = (char)(0xF0 | (cp >> 18)); *o++ = (char)(0x80 | ((cp >> 12) & 0x3F)); *o++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *o++ = (char)(0x80 | (cp & 0x3F)); }
	*out = o;
}

static unsigned hex4(const char *p, bool *ok) {
	unsigned v = 0;
	for (int i = 0; i < 4; ++i) {
		char c = p[i];
		v <<= 4;
		if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
		else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
		else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
		else { *ok = false; return 0; }
	}
	return v;
}

static char *parse_string_raw(struct parser *s) {
	if (*s->p != '"') { s->err = true; return NULL; }
	const char *q = s->p + 1;
	size_t len = 0;
	while (*q && *q != '"') { if (*q == '\\' && q[1]) q++; q++; len++; }
	if (*q != '"') { s->err = true; return NULL; }
	char *out = malloc(len * 4 + 1), *o = out;
	q = s->p + 1;
	while (*q != '"') {
		if (*q != '\\') { *o++ = *q++; continue; }
		q++;
		switch (*q) {
		case 'b': *o++ = '\b'; break; case 'f': *o++ = '\f'; break; case 'n': *o++ = '\n'; break;
		case 'r': *o++ = '\r'; break; case 't': *o++ = '\t'; break;
		case '"': case '\\': case '/': *o++ = *q; break;
		case 'u': {
			bool ok = true;
			unsigned cp = hex4(q + 1, &ok);
			if (!ok) { free(out); s->err = true; return NULL; }
			q += 4;
			if (cp >= 0xD800 && cp <= 0xDBFF && q[1] == '\\' && q[2] == 'u') {
				unsigned lo = hex4(q + 3, &ok);
				if (ok && lo >= 0xDC00 && lo <= 0xDFFF) { cp = 0x10000 + (((cp & 0x3FF) << 10) | (lo & 0x3FF)); q += 6; }
			}
			utf8_put(&o, cp);
			break;
		}
		default: free(out); s->err = true; return NULL;
		}
		q++;
	}
	*o = 0;
	s->p = q + 1;
	return out;
}

static struct crj *parse_number(struct parser *s) {
	char *end = NULL;
	double d = strtod(s->p, &end);                    /* cJSON.c parse_number: strtod on the number text */
	if (end == s->p) { s->err = true; return NULL; }
	struct crj *n = node_new(CRJ_NUMBER);
	n->num = d;
	n->inum = d >= INT_MAX ? INT_MAX : (d <= (double)INT_MIN ? INT_MIN : (int)d);
	s->p = end;
	return n;
}

static struct crj *parse_container(struct parser *s, int depth, bool object) {
	struct crj *n = node_new(object ? CRJ_OBJECT : CRJ_ARRAY), *tail = NULL;
	const char close = object ? '}' : ']';
	s->p++;
	skip_ws(s);
	if (*s->p == close) { s->p++; return n; }
	while (!s->err) {
		skip_ws(s);
		char *key = NULL;
		if (object) {
			key = parse_string_raw(s);
			if (s->err) break;
			skip_ws(s);
			if (*s->p != ':') { free(key); s->err = true; break; }
			s->p++;
		}
		struct crj *v = parse_value(s, depth + 1);
		if (!v) { free(key); s->err = true; break; }
		v->key = key;
		if (tail) tail->next = v; else n->child = v;
		tail = v;
		skip_ws(s);
		if (*s->p == ',') { s->p++; continue; }
		if (*s->p == close) { s->p++; return n; }
		s->err = true;
	}
	crj_free(n);
	return NULL;
}

static struct crj *parse_value(struct parser *s, int depth) {
	if (depth > 512) { s->err = true; return NULL; }
	skip_ws(s);
	if (!strncmp(s->p, "null", 4)) { s->p += 4; return node_new(CRJ_NULL); }
	if (!strncmp(s->p, "false", 5)) { s->p += 5; return node_new(CRJ_FALSE); }
	if (!strncmp(s->p, "true", 4)) { s->p += 4; return node_new(CRJ_TRUE); }
	if (*s->p == '"') { char *str = parse_string_raw(s); if (s->err) return NULL; struct crj *n = node_new(CRJ_STRING); n->str = str; return n; }
	if (*s->p == '-' || (*s->p >= '0' && *s->p <= '9')) return parse_number(s);
	if (*s->p == '[') return parse_container(s, depth, false);
	if (*s->p == '{') return parse_container(s, depth, true);
	s->err = true;
	return NULL;
}

struct crj *crj_parse(const char *text) {
	if (!text) return NULL;
	struct parser s = { text, false };
	if ((unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF) s.p += 3;
	struct crj *root = parse_value(&s, 0);
	if (s.err) { crj_free(root); return NULL; }
	return root;
}

static int ci_cmp(const char *a, const char *b) {
	for (; tolower((unsigned char)*a) == tolower((unsigned char)*b); ++a, ++b) if (!*a) return 0;
	return 1;
}

const struct crj *crj_get(const struct crj *obj, const char *key) {
	if (!obj || !key || (obj->type != CRJ_OBJECT && obj->type != CRJ_ARRAY)) return NULL;
	for (const struct crj *c = obj->child; c; c = c->next) if (c->key && ci_cmp(c->key, key) == 0) return c;
	return NULL;
}
const struct crj *crj_at(const struct crj *arr, int index) {
	if (!arr || index < 0) return NULL;
	const struct crj *c = arr->child;
	while (c && index-- > 0) c = c->next;
	return c;
}
int crj_size(const struct crj *arr) {
	int n = 0;
	if (arr) for (const struct crj *c = arr->child; c; c = c->next) n++;
	return n;
}
